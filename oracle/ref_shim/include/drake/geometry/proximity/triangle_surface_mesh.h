// SHIM (test infrastructure, ours): the accessors of drake::geometry::TriangleSurfaceMesh<double> the reference's
// ray caster calls (bvh.cpp:92-105): num_elements(), element(i).vertex(k), vertex(v).  Plain storage, no arithmetic.
#pragma once
#include <drake/common/eigen_types.h>

#include <utility>
#include <vector>

namespace drake {
namespace geometry {

class SurfaceTriangle {
public:
	SurfaceTriangle(int a, int b, int c) : v_{ a, b, c } {}
	int vertex(int i) const { return v_[i]; }

private:
	int v_[3];
};

template <typename T>
class TriangleSurfaceMesh {
public:
	static constexpr int kVertexPerElement = 3;
	TriangleSurfaceMesh(std::vector<SurfaceTriangle> tris, std::vector<Vector3<T>> verts)
	    : tris_(std::move(tris)), verts_(std::move(verts))
	{
	}
	int num_elements() const { return static_cast<int>(tris_.size()); }
	int num_triangles() const { return num_elements(); }
	int num_vertices() const { return static_cast<int>(verts_.size()); }
	const SurfaceTriangle &element(int i) const { return tris_[i]; }
	const Vector3<T> &vertex(int v) const { return verts_[v]; }

private:
	std::vector<SurfaceTriangle> tris_;
	std::vector<Vector3<T>> verts_;
};

} // namespace geometry
} // namespace drake
