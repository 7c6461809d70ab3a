// SHIM (test infrastructure, ours): the two accessors of drake::geometry::ContactSurface<double> the reference's
// sensors call: tri_mesh_W() (bvh.cpp:92-105) and tri_e_MN() (flat_tactile_sensor.cpp:350).  A container only.
#pragma once
#include <drake/geometry/proximity/triangle_surface_mesh.h>
#include <drake/geometry/proximity/triangle_surface_mesh_field.h>

#include <memory>

namespace drake {
namespace geometry {

template <typename T>
class ContactSurface {
public:
	ContactSurface(std::unique_ptr<TriangleSurfaceMesh<T>> mesh_W,
	               std::unique_ptr<TriangleSurfaceMeshFieldLinear<T, T>> e_MN)
	    : mesh_W_(std::move(mesh_W)), e_MN_(std::move(e_MN))
	{
	}
	const TriangleSurfaceMesh<T> &tri_mesh_W() const { return *mesh_W_; }
	const TriangleSurfaceMeshFieldLinear<T, T> &tri_e_MN() const { return *e_MN_; }

private:
	std::unique_ptr<TriangleSurfaceMesh<T>> mesh_W_;
	std::unique_ptr<TriangleSurfaceMeshFieldLinear<T, T>> e_MN_;
};

} // namespace geometry
} // namespace drake
