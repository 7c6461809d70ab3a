// Host-side mesh TOPOLOGY + vertex generation for the per-geom hydroelastic representation
// (north star stage 1).  Replaces the Drake Make*VolumeMesh / Make*SurfaceMesh / Make*PressureField
// calls of mujoco_contact_surfaces_plugin.cpp:650-797.  Per-element derived data (pressure
// gradients, tet half spaces, triangle normals, LBVH) is built on the GPU (kernels_build.cu).
// The enumeration order of vertices and elements is specified in DESIGN.md "Mesh specification".
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace hcs {

struct HostMesh {
	bool soft = false;
	bool plane = false;
	std::vector<double> verts;    // xyz triples, geom frame
	std::vector<int32_t> elems;   // 4 per tet (soft) or 3 per triangle (rigid)
	std::vector<double> pressure; // per vertex (soft only)
	int n_verts() const { return (int)verts.size() / 3; }
	int n_elems() const { return soft ? (int)elems.size() / 4 : (int)elems.size() / 3; }
};

// mj_type / size / props as in hcs_add_geom.  Returns false and sets err for unsupported input.
bool build_geom_mesh(int mj_type, const double size[3], const float *mesh_vert, int n_vert, const int32_t *mesh_face,
                     int n_face, const double props[5], HostMesh &out, std::string &err);

// Sphere / ellipsoid geoms are a unit sphere (single interior vertex, boundary refined `level` times) scaled by the
// semi-axes; the level follows from the sizes and the resolution hint.  Exposed for per-environment sizes
// (hcs_set_env_sizes): one unit mesh, every environment's vertices and pressures are generated from it on the GPU.
int sphere_like_level(int mj_type, const double size[3], double hint);
// vertices of the unit sphere mesh of that level in generation order (vertex 0 = centre)
void unit_sphere_vertices(int level, std::vector<double> &verts);

} // namespace hcs
