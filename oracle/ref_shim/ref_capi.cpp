// TEST INFRASTRUCTURE (ours): C-ABI harness around the REFERENCE's own float32 ray caster.
//
// This file is compiled TOGETHER with the unmodified reference source
//   /root/reference/mujoco_contact_surface_sensors/src/bvh.cpp            (BVH, TLAS, IntersectTriangle, slab tests)
//   /root/reference/mujoco_contact_surface_sensors/include/.../bvh.h, float3.h
// against the shim headers in oracle/ref_shim/include (container-only stand-ins for the four Drake accessors,
// Eigen::Map / cast<float> and boost::shared_ptr the file touches) into oracle/_ref/libref_bvh*.so by
// oracle/ref_shim/Makefile.  Nothing of the reference is copied into this repository: the sources are compiled where
// they lie.  The harness drives the ray caster the way FlatTactileSensor::bvh_update does
// (flat_tactile_sensor.cpp:285-302: `BVH bvh[n]; bvh[i] = BVH(gc->s); TLAS tlas(bvh, n); tlas.build();` and
// :329-337: `Ray ray; ray.O = ...; ray.D = ...; ray.hit.t = 1e30f; tlas.intersect(ray);`).
#include <mujoco_contact_surface_sensors/bvh.h>

#include <cstdint>
#include <vector>

using namespace mujoco_ros::contact_surfaces::sensors;
using drake::geometry::ContactSurface;
using drake::geometry::SurfaceTriangle;
using drake::geometry::TriangleSurfaceMesh;
using drake::geometry::TriangleSurfaceMeshFieldLinear;

namespace {
struct RefScene {
	std::vector<std::shared_ptr<ContactSurface<double>>> surfaces;
	BVH *bvh   = nullptr; // array, as the VLA in bvh_update
	TLAS *tlas = nullptr;
	~RefScene()
	{
		delete tlas;
		delete[] bvh;
	}
};
} // namespace

extern "C" {

// 1 when the reference was compiled with -DUSE_SSE (its CMake default when SSE4 is present), else 0
int ref_use_sse()
{
#ifdef USE_SSE
	return 1;
#else
	return 0;
#endif
}

// n_surf contact surfaces as triangle soups: surface s has n_tri[s] triangles; verts = all triangles back to back,
// 9 doubles each (world frame, the doubles Drake's tri_mesh_W() would hold); press = 3 vertex pressures each.
void *ref_tlas_create(int n_surf, const int *n_tri, const double *verts, const double *press)
{
	RefScene *sc = new RefScene;
	sc->bvh      = new BVH[n_surf];
	size_t off   = 0;
	for (int s = 0; s < n_surf; ++s) {
		std::vector<SurfaceTriangle> tris;
		std::vector<drake::Vector3<double>> v;
		std::vector<double> e;
		for (int t = 0; t < n_tri[s]; ++t, ++off) {
			for (int k = 0; k < 3; ++k) {
				v.emplace_back(verts[9 * off + 3 * k], verts[9 * off + 3 * k + 1], verts[9 * off + 3 * k + 2]);
				e.push_back(press ? press[3 * off + k] : 0.0);
			}
			tris.emplace_back(3 * t, 3 * t + 1, 3 * t + 2);
		}
		auto mesh  = std::make_unique<TriangleSurfaceMesh<double>>(std::move(tris), std::move(v));
		auto field = std::make_unique<TriangleSurfaceMeshFieldLinear<double, double>>(std::move(e), mesh.get());
		sc->surfaces.push_back(std::make_shared<ContactSurface<double>>(std::move(mesh), std::move(field)));
		sc->bvh[s] = BVH(sc->surfaces.back()); // flat_tactile_sensor.cpp:288
	}
	sc->tlas = new TLAS(sc->bvh, n_surf); // :301-302
	sc->tlas->build();
	return sc;
}

// casts n rays; O, D: 3 floats per ray; tuv: 3 floats per ray out (t = 1e30f on a miss); id: (blas << 20) + triangle
void ref_tlas_cast(void *h, int n, const float *O, const float *D, float *tuv, uint32_t *id)
{
	RefScene *sc = (RefScene *)h;
	for (int i = 0; i < n; ++i) {
		Ray ray; // flat_tactile_sensor.cpp:332-337
		ray.d0.data.O        = float3(O[3 * i], O[3 * i + 1], O[3 * i + 2]);
		ray.d1.data.D        = float3(D[3 * i], D[3 * i + 1], D[3 * i + 2]);
		ray.hit.t            = 1e30f;
		ray.hit.u            = 0;
		ray.hit.v            = 0;
		ray.hit.bvh_triangle = 0;
		sc->tlas->intersect(ray);
		tuv[3 * i]     = ray.hit.t;
		tuv[3 * i + 1] = ray.hit.u;
		tuv[3 * i + 2] = ray.hit.v;
		id[i]          = ray.hit.bvh_triangle;
	}
}

// flat_tactile_sensor.cpp:350: tlas.blas[blas_idx].surface->tri_e_MN().Evaluate(tri_idx, bary)
double ref_evaluate(void *h, uint32_t blas_idx, uint32_t tri_idx, const double *bary)
{
	RefScene *sc = (RefScene *)h;
	return sc->tlas->blas[blas_idx].surface->tri_e_MN().Evaluate(tri_idx, Eigen::Vector3d(bary[0], bary[1], bary[2]));
}

void ref_tlas_destroy(void *h) { delete (RefScene *)h; }

// single primitives, for known-answer tests of the restatement
void ref_intersect_triangle(const float *O, const float *D, const float *v0, const float *v1, const float *v2,
                            float t_in, float *tuv_out, int *hit_out)
{
	Ray ray;
	ray.d0.data.O        = float3(O[0], O[1], O[2]);
	ray.d1.data.D        = float3(D[0], D[1], D[2]);
	ray.hit.t            = t_in;
	ray.hit.u            = 0;
	ray.hit.v            = 0;
	ray.hit.bvh_triangle = 0xffffffffu;
	Triangle tri;
	tri.vertex0 = float3(v0[0], v0[1], v0[2]);
	tri.vertex1 = float3(v1[0], v1[1], v1[2]);
	tri.vertex2 = float3(v2[0], v2[1], v2[2]);
	IntersectTriangle(ray, tri, 7u);
	tuv_out[0] = ray.hit.t, tuv_out[1] = ray.hit.u, tuv_out[2] = ray.hit.v;
	*hit_out = ray.hit.bvh_triangle == 7u;
}

float ref_intersect_aabb(const float *O, const float *D, float t_in, const float *bmin, const float *bmax)
{
	Ray ray;
	ray.d0.data.O  = float3(O[0], O[1], O[2]);
	ray.d1.data.D  = float3(D[0], D[1], D[2]);
	ray.d2.data.rD = float3(1.0f / D[0], 1.0f / D[1], 1.0f / D[2]);
	ray.hit.t      = t_in;
	return IntersectAABB(ray, float3(bmin[0], bmin[1], bmin[2]), float3(bmax[0], bmax[1], bmax[2]));
}

} // extern "C"
