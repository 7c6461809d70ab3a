// TEST INFRASTRUCTURE (ours): forward declaration of the reference's IntersectTriangle (defined with external
// linkage at mujoco_contact_surface_sensors/src/bvh.cpp:49, not declared in bvh.h) so that the harness can call it.
#pragma once
namespace mujoco_ros { namespace contact_surfaces { namespace sensors {
struct Ray;
struct Triangle;
void IntersectTriangle(Ray &ray, const Triangle &tri, const unsigned int bvh_triangle);
}}}
