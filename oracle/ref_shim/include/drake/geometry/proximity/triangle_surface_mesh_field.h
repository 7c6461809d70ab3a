// SHIM (test infrastructure, ours): TriangleSurfaceMeshFieldLinear<double,double>::Evaluate(element, barycentric)
// as the reference's flat sensor calls it (flat_tactile_sensor.cpp:350).  Restates Drake v1.8.0
// geometry/proximity/mesh_field_linear.h Evaluate(): value = b[0]*v0; value += b[1]*v1; value += b[2]*v2.
// This one function is Drake arithmetic and therefore part of the UNPINNED share of the oracle.
#pragma once
#include <drake/geometry/proximity/triangle_surface_mesh.h>

namespace drake {
namespace geometry {

template <typename FieldValue, typename T>
class TriangleSurfaceMeshFieldLinear {
public:
	TriangleSurfaceMeshFieldLinear(std::vector<FieldValue> values, const TriangleSurfaceMesh<T> *mesh)
	    : values_(std::move(values)), mesh_(mesh)
	{
	}
	FieldValue Evaluate(int e, const Eigen::Vector3d &b) const
	{
		const SurfaceTriangle &el = mesh_->element(e);
		FieldValue value          = b[0] * values_[el.vertex(0)];
		for (int i = 1; i < 3; ++i)
			value += b[i] * values_[el.vertex(i)];
		return value;
	}

private:
	std::vector<FieldValue> values_;
	const TriangleSurfaceMesh<T> *mesh_;
};

} // namespace geometry
} // namespace drake
