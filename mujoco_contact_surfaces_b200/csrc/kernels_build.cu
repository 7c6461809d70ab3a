// K1 build_fields — per-element derived data, built once at hcs_finalize (and hcs_update_geom).
// Replaces the Drake VolumeMeshFieldLinear constructor (CalcGradBarycentric, value at mesh origin)
// and TriangleSurfaceMesh face normals the reference gets at plugin.cpp:654-662, and hoists the
// per-candidate half-space construction of mesh_intersection.cc ClipTriangleByTetrahedron
// (normal = (B-A)x(C-A), normalized; d = nhat.A) out of the step: those values depend on the tet
// only, so precomputing them is bit-identical to recomputing them per candidate.
// Streaming kernel: one thread per element, 128..192-B records written with full-sector stores.
#include "dmath.cuh"
#include "hcs_internal.h"

namespace hcs {

__global__ void __launch_bounds__(128) build_tets_kernel(GeomDev g)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= g.n_elems)
		return;
	int4 idx = reinterpret_cast<const int4 *>(g.elems)[t];
	int vi[4] = { idx.x, idx.y, idx.z, idx.w };
	D3 v[4];
	double e[4];
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		v[i] = ld3(g.verts + 3 * (size_t)vi[i]);
		e[i] = g.pressure[vi[i]];
	}
	TetGeom tg;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		tg.v[i][0] = v[i].x, tg.v[i][1] = v[i].y, tg.v[i][2] = v[i].z;
		tg.e[i] = e[i];
	}
	g.tet_geom[t] = tg;

	TetField tf;
	// gradient: sum_i e_i * grad(b_i), grad(b_i) = (AB x AC) / ((AB x AC) . AV), A,B,C = vertices i+1,i+2,i+3
	D3 grad = mk(0, 0, 0);
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		D3 V = v[i], A = v[(i + 1) & 3], B = v[(i + 2) & 3], C = v[(i + 3) & 3];
		D3 area_vec = cross(B - A, C - A);
		double sv   = dot(area_vec, V - A);
		D3 gb       = area_vec / sv;
		grad        = i == 0 ? e[0] * gb : grad + e[i] * gb;
	}
	tf.grad[0] = grad.x, tf.grad[1] = grad.y, tf.grad[2] = grad.z;
	tf.e0 = e[0] - dot(grad, v[0]);
	D3 gh = normalized(grad);
	tf.ghat[0] = gh.x, tf.ghat[1] = gh.y, tf.ghat[2] = gh.z;
	tf.pad = 0;
	const int F[4][3] = { { 1, 2, 3 }, { 0, 3, 2 }, { 0, 1, 3 }, { 0, 2, 1 } };
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		D3 A = v[F[k][0]], B = v[F[k][1]], C = v[F[k][2]];
		D3 n = normalized(cross(B - A, C - A));
		tf.plane[k][0] = n.x, tf.plane[k][1] = n.y, tf.plane[k][2] = n.z;
		tf.plane[k][3] = dot(n, A);
	}
	g.tet_field[t] = tf;

	// float copy for the conservative leaf filter of the broadphase (kernels_broadphase.cu): it only rejects pairs
	// the exact fp64 tests reject with a margin far above the rounding of these values
	TetLeaf32 tl;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
#pragma unroll
		for (int a = 0; a < 4; ++a)
			tl.plane[k][a] = (float)tf.plane[k][a];
#pragma unroll
		for (int a = 0; a < 3; ++a)
			tl.v[k][a] = (float)tg.v[k][a];
	}
	tl.ghat[0] = (float)gh.x, tl.ghat[1] = (float)gh.y, tl.ghat[2] = (float)gh.z;
	tl.pad = 0.f;
	g.tet_leaf32[t] = tl;
	TetLeafSS32 ts;
	ts.grad[0] = (float)grad.x, ts.grad[1] = (float)grad.y, ts.grad[2] = (float)grad.z, ts.e0 = (float)tf.e0;
	ts.ghat[0] = tl.ghat[0], ts.ghat[1] = tl.ghat[1], ts.ghat[2] = tl.ghat[2], ts.pad = 0.f;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		ts.pad2[k] = 0.f;
#pragma unroll
		for (int a = 0; a < 3; ++a)
			ts.v[k][a] = tl.v[k][a];
	}
	g.tet_leafss32[t] = ts;
	TetBox32 tb;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const double lo = fmin(fmin(tg.v[0][a], tg.v[1][a]), fmin(tg.v[2][a], tg.v[3][a]));
		const double hi = fmax(fmax(tg.v[0][a], tg.v[1][a]), fmax(tg.v[2][a], tg.v[3][a]));
		tb.lo[a] = __double2float_rd(lo), tb.hi[a] = __double2float_ru(hi);
	}
	tb.pad[0] = tb.pad[1] = 0.f;
	g.tet_box32[t] = tb;
}

__global__ void __launch_bounds__(128) build_tris_kernel(GeomDev g)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= g.n_elems)
		return;
	TriRec r;
	D3 v[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		v[i]      = ld3(g.verts + 3 * (size_t)g.elems[3 * (size_t)t + i]);
		r.v[i][0] = v[i].x, r.v[i][1] = v[i].y, r.v[i][2] = v[i].z;
	}
	D3 cr     = cross(v[1] - v[0], v[2] - v[0]);
	double nn = sqrt(dot(cr, cr));
	D3 n      = nn != 0.0 ? cr / nn : cr;
	r.n[0] = n.x, r.n[1] = n.y, r.n[2] = n.z;
	g.tris[t] = r;
}

// Per-environment sizes of sphere / ellipsoid geoms (hcs_set_env_sizes): vertices = unit-sphere vertex x semi-axes and the
// pressure extent field of mesh_host.cpp build_geom_mesh, same expressions in the same order (IEEE sqrt / division, no
// FMA contraction: bit-identical to the host generator).  One thread per (environment, vertex).
// vol_offset: 1 for rigid geoms, whose surface mesh is the volume mesh without its centre vertex (vertex 0).
__global__ void __launch_bounds__(128) sphere_env_verts_kernel(const double *unit, int n_verts, int vol_offset, const double *sizes,
                                                               int n_env, int is_sphere, double E, double *verts, double *pressure)
{
	const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long)n_env * n_verts)
		return;
	const int env = (int)(i / n_verts), v = (int)(i - (long)env * n_verts);
	const double a = sizes[3 * env], b = is_sphere ? a : sizes[3 * env + 1], c = is_sphere ? a : sizes[3 * env + 2];
	const double *u = unit + 3 * (size_t)(v + vol_offset);
	const D3 p      = mk(u[0] * a, u[1] * b, u[2] * c);
	verts[3 * (size_t)i] = p.x, verts[3 * (size_t)i + 1] = p.y, verts[3 * (size_t)i + 2] = p.z;
	if (pressure) {
		const D3 q       = is_sphere ? p : mk(p.x / a, p.y / b, p.z / c);
		const double rad = sqrt(dot(q, q));
		double ext       = is_sphere ? 1.0 - rad / a : 1.0 - rad;
		if (fabs(ext) < 1e-14)
			ext = 0.0;
		pressure[i] = E * ext;
	}
}

void launch_sphere_env_verts(const double *unit, int n_verts, int vol_offset, const double *sizes, int n_env, int is_sphere, double E,
                             double *verts, double *pressure, cudaStream_t s)
{
	const long n = (long)n_env * n_verts;
	if (n > 0)
		sphere_env_verts_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(unit, n_verts, vol_offset, sizes, n_env, is_sphere, E, verts,
		                                                                  pressure);
}

void launch_build_tets(const GeomDev &g, cudaStream_t s)
{
	if (g.n_elems > 0)
		build_tets_kernel<<<(g.n_elems + 127) / 128, 128, 0, s>>>(g);
}
void launch_build_tris(const GeomDev &g, cudaStream_t s)
{
	if (g.n_elems > 0)
		build_tris_kernel<<<(g.n_elems + 127) / 128, 128, 0, s>>>(g);
}

} // namespace hcs
