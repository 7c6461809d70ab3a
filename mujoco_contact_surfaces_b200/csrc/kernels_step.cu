// Per-step kernels of the hydroelastic contact engine (sm_100a, fp64 geometry mode, -fmad=false).
//
//   K3 broadphase_kernel      one warp per (env, pair, query slice): each lane walks the soft geom's LBVH
//                             with one query element of the other geom; leaf hits are compacted with a
//                             warp ballot + popc prefix into the warp's candidate slab (deterministic order).
//                             Replaces Bvh<Obb,.>::Collide inside the Drake queries called at
//                             mujoco_contact_surfaces_plugin.cpp:284-303.
//   K4 narrow_tet_tri_kernel  one thread per (tet, triangle) candidate: normal/gradient cull, Sutherland-
//                             Hodgman clip against the tet's four precomputed half spaces, duplicate removal,
//                             polygon quadrature (plugin.cpp:320-409) and the force law (plugin.cpp:411-483).
//                             Restates mesh_intersection.cc.
//   K5 narrow_tet_plane_kernel one thread per (tet, half space): marching-tets slice (App. A.5).
//   K6 narrow_tet_tet_kernel  one thread per (tet, tet) candidate: equal-pressure plane, slice + clip (A.6).
//   K7 finalize kernels       fixed-order reduction of the per-warp partial sums to per-pair wrenches and
//                             per-geom wrenches (replaces two mj_applyFT per face, plugin.cpp:477-482).
//
// The polygon vertex arithmetic follows the oracle's (= restated Drake) operation order exactly; the
// quadrature uses the known unit normal instead of per-fan-triangle norms (differences ~1e-16 relative).
#include "dmath.cuh"
#include "hcs_internal.h"

namespace hcs {

#define FULL_MASK 0xffffffffu
constexpr int MAXV       = 8;
constexpr int STACK_SIZE = 64;
constexpr int WARPS_PER_BLOCK = 4;
constexpr int BLOCK           = 32 * WARPS_PER_BLOCK;

struct WarpCtx { // warp-uniform per (env, pair) data
	Xform X_WA;    // soft geom A (computation frame) -> world
	D3 xA, wA, vA; // origin, angular, linear velocity of geom A (world)
	D3 xB, wB, vB;
	double dissipation, mu;
	int apply;
};

struct Acc {
	D3 F, tau, ac;
	double area;
	int n_polygons, n_faces, n_points, n_candidates;
};

__device__ __forceinline__ void load_vel(const double *vel, int n_geoms, int env, int g, D3 &w, D3 &v)
{
	const double *p = vel + ((size_t)env * n_geoms + g) * 6;
	w               = ld3(p);
	v               = ld3(p + 3);
}

__device__ __forceinline__ WarpCtx make_ctx(const PairDesc &P, const StepIO &io, int env, const Xform &X_WA,
                                            const Xform &X_WB)
{
	WarpCtx c;
	c.X_WA = X_WA;
	c.xA   = X_WA.p;
	c.xB   = X_WB.p;
	load_vel(io.vel, io.n_geoms, env, P.gA, c.wA, c.vA);
	load_vel(io.vel, io.n_geoms, env, P.gB, c.wB, c.vB);
	c.dissipation = P.dissipation;
	c.mu          = P.mu;
	c.apply       = io.apply_forces;
	return c;
}

// passiveCallback force law (plugin.cpp:440-475) for one quadrature point; A = M (unswapped labelling)
__device__ __forceinline__ D3 face_force(D3 p, D3 n, double fn0, double k, const WarpCtx &c)
{
	D3 vAq    = c.vA + cross(c.wA, p - c.xA);
	D3 vBq    = c.vB + cross(c.wB, p - c.xB);
	D3 vrel   = vAq - vBq;
	double vn = dot(vrel, n);
	double fn = fmax(0., 1. - c.dissipation * vn) * (fn0 - 0.001 * k * vn);
	if (!c.apply)
		return mk(0, 0, 0);
	D3 vt        = vrel - n * vn;
	double eps   = 1.0e-4 * 1.0e-2;
	eps          = eps * eps;
	double vslip = sqrt(dot(vt, vt) + eps);
	D3 that      = vt / vslip;
	double mu_r  = c.mu;
	double s     = vslip / 1.0e-4;
	if (s < 1)
		mu_r = c.mu * s * (2.0 - s);
	D3 fslip = -mu_r * that * fn;
	return fslip + fn * n;
}

// ---- Sutherland-Hodgman step: ClipPolygonByHalfSpace + CalcIntersection (mesh_intersection.cc) ----
__device__ __forceinline__ int clip_halfspace(const D3 *in, int n, D3 nh, double d, D3 *out)
{
	double sd[MAXV];
	for (int i = 0; i < n; ++i)
		sd[i] = dot(nh, in[i]) - d;
	int m = 0;
	for (int i = 0; i < n; ++i) {
		int ip   = i == 0 ? n - 1 : i - 1;
		bool cin = sd[i] <= 0, pin = sd[ip] <= 0;
		if (cin != pin) { // CalcIntersection(current, previous)
			double a = sd[i], b = sd[ip];
			double wa = b / (b - a);
			double wb = 1.0 - wa;
			out[m++]  = wa * in[i] + wb * in[ip];
		}
		if (cin)
			out[m++] = in[i];
	}
	return m;
}

// RemoveDuplicateVertices: std::unique over consecutive near vertices, then last vs first
__device__ __forceinline__ int remove_duplicates(D3 *p, int n)
{
	const double eps2 = 1e-14 * 1e-14;
	if (n == 0)
		return 0;
	int m = 1;
	for (int i = 1; i < n; ++i) {
		D3 d = p[m - 1] - p[i];
		if (!(dot(d, d) < eps2))
			p[m++] = p[i];
	}
	if (m >= 3) {
		D3 d = p[0] - p[m - 1];
		if (dot(d, d) < eps2)
			--m;
	}
	return m;
}

__constant__ int c_tet_edges[6][2]     = { { 0, 1 }, { 1, 2 }, { 2, 0 }, { 0, 3 }, { 1, 3 }, { 2, 3 } };
__constant__ int c_marching_tets[16][4] = { { -1, -1, -1, -1 }, { 0, 3, 2, -1 }, { 0, 1, 4, -1 }, { 4, 3, 2, 1 },
	                                        { 1, 2, 5, -1 },    { 0, 3, 5, 1 },  { 0, 2, 5, 4 },  { 3, 5, 4, -1 },
	                                        { 3, 4, 5, -1 },    { 4, 5, 2, 0 },  { 1, 5, 3, 0 },  { 1, 5, 2, -1 },
	                                        { 1, 2, 3, 4 },     { 0, 4, 1, -1 }, { 0, 2, 3, -1 }, { -1, -1, -1, -1 } };

struct EmitInfo { // provenance for the optional dumps
	int env, pair, elemA, elemB, slot;
};

// Quadrature + force accumulation of one contact polygon.
//   P[0..n): vertices in the builder frame (A's frame, or world when IDENT), winding such that the
//   right-handed normal is nhat (unit, points into A); e[i]: pressures; grad: sampled-field gradient
//   (builder frame); gN: -grad_N . nhat or +inf.  TRI selects kTriangle (centroid fan) vs kPolygon.
template <bool TRI, bool IDENT>
__device__ __forceinline__ void integrate_polygon(const D3 *P, int n, D3 nhat, D3 grad, const double *e, double gN,
                                                  const WarpCtx &c, const PairDesc &Pd, const StepIO &io,
                                                  const EmitInfo &info, Acc &acc, int &tri_faces, D3 *W_out, D3 &cW_out,
                                                  double &ec_out)
{
	const double kInf = __longlong_as_double(0x7ff0000000000000LL);
	double gM         = dot(grad, nhat);
	D3 nW             = IDENT ? nhat : rot(c.X_WA.R, nhat);
	acc.n_polygons += 1;
	tri_faces = 0;
	// polygon centroid (contact_surface_utility.cc CalcPolygonCentroid): fan about vertex 0, signed
	// areas measured along nhat
	D3 p0     = P[0];
	double A2 = 0;
	D3 csum   = mk(0, 0, 0);
	for (int i = 1; i < n - 1; ++i) {
		double a2 = dot(cross(P[i] - p0, P[i + 1] - p0), nhat);
		A2 += a2;
		csum = csum + a2 * ((p0 + P[i]) + P[i + 1]);
	}
	D3 cen;
	if (n == 3)
		cen = ((P[0] + P[1]) + P[2]) / 3.0;
	else
		cen = A2 != 0.0 ? csum / (3.0 * A2) : p0;
	// A face whose winding opposes nhat (only possible for a negatively oriented tet of a user mesh) gets
	// the flipped normal, like the mesh constructors that derive face normals from the winding.
	if (!TRI) {
		acc.n_faces += 1;
		double sg   = A2 < 0 ? -1.0 : 1.0;
		double area = 0.5 * (sg * A2);
		double gMf = sg * gM, gNf = gN == kInf ? gN : sg * gN;
		bool g_ok  = !(gMf < 1.0e-14 || gNf < 1.0e-14);
		double g   = 1.0 / (1.0 / gMf + 1.0 / gNf);
		nW         = sg * nW;
		D3 cW      = IDENT ? cen : apply(c.X_WA, cen);
		if (area > 0) {
			acc.area += area;
			acc.ac = acc.ac + area * cW;
		}
		if (area > 1.0e-14 && g_ok) {
			double pc  = e[0] + dot(grad, cen - p0);
			double fn0 = area * pc, k = area * g;
			D3 f       = face_force(cW, nW, fn0, k, c);
			acc.F      = acc.F + f;
			acc.tau    = acc.tau + cross(cW, f);
			acc.n_points += 1;
			if (io.max_faces > 0) {
				int slot = atomicAdd(io.face_count, 1);
				if (slot < io.max_faces) {
					hcs_face &o = io.faces[slot];
										o.p[0] = cW.x, o.p[1] = cW.y, o.p[2] = cW.z;
					o.n[0] = Pd.sign * nW.x, o.n[1] = Pd.sign * nW.y, o.n[2] = Pd.sign * nW.z;
					o.fn0 = fn0, o.stiffness = k, o.damping = c.dissipation;
					o.f[0] = Pd.sign * f.x, o.f[1] = Pd.sign * f.y, o.f[2] = Pd.sign * f.z;
					o.env = info.env, o.pair = info.pair;
					o.elemM = Pd.sign > 0 ? info.elemA : info.elemB;
					o.elemN = Pd.sign > 0 ? info.elemB : info.elemA;
					o.nverts = n, o.face = 0;
				}
			}
		}
		return;
	}
	// kTriangle: TriMeshBuilder::AddPolygon — centroid vertex, pressure by the gradient, fan (prev,next,c)
	double ec = e[0] + dot(grad, cen - p0);
	D3 cW     = IDENT ? cen : apply(c.X_WA, cen);
	for (int i = 0; i < n; ++i)
		W_out[i] = IDENT ? P[i] : apply(c.X_WA, P[i]);
	cW_out    = cW;
	ec_out    = ec;
	tri_faces = n;
	acc.n_faces += n;
	int cur = n - 1;
	for (int i = 0; i < n; ++i) {
		D3 a = P[cur], b = P[i];
		double a2   = dot(cross(b - a, cen - a), nhat);
		double sg   = a2 < 0 ? -1.0 : 1.0;
		double area = 0.5 * (sg * a2);
		double gMf = sg * gM, gNf = gN == kInf ? gN : sg * gN;
		bool g_ok  = !(gMf < 1.0e-14 || gNf < 1.0e-14);
		double g   = 1.0 / (1.0 / gMf + 1.0 / gNf);
		D3 nWf     = sg * nW;
		D3 fc      = ((W_out[cur] + W_out[i]) + cW) / 3.0;
		if (area > 0) {
			acc.area += area;
			acc.ac = acc.ac + area * fc;
		}
		if (area > 1.0e-14 && g_ok) {
			double b3 = 1 / 3.;
			double pc = b3 * e[cur];
			pc += b3 * e[i];
			pc += b3 * ec;
			double fn0 = area * pc, k = area * g;
			D3 f       = face_force(fc, nWf, fn0, k, c);
			acc.F      = acc.F + f;
			acc.tau    = acc.tau + cross(fc, f);
			acc.n_points += 1;
			if (io.max_faces > 0) {
				int slot = atomicAdd(io.face_count, 1);
				if (slot < io.max_faces) {
					hcs_face &o = io.faces[slot];
										o.p[0] = fc.x, o.p[1] = fc.y, o.p[2] = fc.z;
					o.n[0] = Pd.sign * nWf.x, o.n[1] = Pd.sign * nWf.y, o.n[2] = Pd.sign * nWf.z;
					o.fn0 = fn0, o.stiffness = k, o.damping = c.dissipation;
					o.f[0] = Pd.sign * f.x, o.f[1] = Pd.sign * f.y, o.f[2] = Pd.sign * f.z;
					o.env = info.env, o.pair = info.pair;
					o.elemM = Pd.sign > 0 ? info.elemA : info.elemB;
					o.elemN = Pd.sign > 0 ? info.elemB : info.elemA;
					o.nverts = n, o.face = i;
				}
			}
		}
		cur = i;
	}
}

// Warp-cooperative append of this lane's fan triangles to the tactile pool (ballot-free exclusive scan
// over lane counts, one atomicAdd per warp).
__device__ __forceinline__ void emit_tactile(int n_faces, const D3 *W, D3 cW, const double *e, double ec,
                                             const PairDesc &Pd, const StepIO &io, const EmitInfo &info, int lane)
{
	int incl = n_faces;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int v = __shfl_up_sync(FULL_MASK, incl, o);
		if (lane >= o)
			incl += v;
	}
	int total = __shfl_sync(FULL_MASK, incl, 31);
	if (total == 0)
		return;
	int base = 0;
	if (lane == 0)
		base = atomicAdd(io.tri_count, total);
	base    = __shfl_sync(FULL_MASK, base, 0);
	int pos = base + incl - n_faces;
	if (n_faces > 0) {
		int cur = n_faces - 1;
		for (int i = 0; i < n_faces; ++i, ++pos) {
			if (pos < io.max_tris) {
				// (prev, next, centroid); the M/N swap of ContactSurface reverses winding by swapping the
				// first two vertices
				int ia = Pd.sign > 0 ? cur : i, ib = Pd.sign > 0 ? i : cur;
				TactileTri t;
				t.v[0] = (float)W[ia].x, t.v[1] = (float)W[ia].y, t.v[2] = (float)W[ia].z;
				t.v[3] = (float)W[ib].x, t.v[4] = (float)W[ib].y, t.v[5] = (float)W[ib].z;
				t.v[6] = (float)cW.x, t.v[7] = (float)cW.y, t.v[8] = (float)cW.z;
				t.e[0] = e[ia], t.e[1] = e[ib], t.e[2] = ec;
				t.env = info.env, t.pair = info.pair, t.order = info.slot * 8 + i;
				io.tri_pool[pos] = t;
			} else {
				atomicOr(io.flags, 2);
			}
			cur = i;
		}
	}
}

__device__ __forceinline__ void reduce_and_store(Acc acc, SlicePartial *out, int lane)
{
	double d[10] = { acc.F.x, acc.F.y, acc.F.z, acc.tau.x, acc.tau.y, acc.tau.z, acc.area, acc.ac.x, acc.ac.y, acc.ac.z };
	int n[4]     = { acc.n_polygons, acc.n_faces, acc.n_points, acc.n_candidates };
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
		for (int k = 0; k < 10; ++k)
			d[k] += __shfl_xor_sync(FULL_MASK, d[k], o);
#pragma unroll
		for (int k = 0; k < 4; ++k)
			n[k] += __shfl_xor_sync(FULL_MASK, n[k], o);
	}
	if (lane == 0) {
		SlicePartial sp;
		sp.F[0] = d[0], sp.F[1] = d[1], sp.F[2] = d[2];
		sp.tau[0] = d[3], sp.tau[1] = d[4], sp.tau[2] = d[5];
		sp.area = d[6];
		sp.ac[0] = d[7], sp.ac[1] = d[8], sp.ac[2] = d[9];
		sp.n_polygons = n[0], sp.n_faces = n[1], sp.n_points = n[2], sp.n_candidates = n[3];
		*out = sp;
	}
}

__device__ __forceinline__ Acc zero_acc()
{
	Acc a;
	a.F = a.tau = a.ac = mk(0, 0, 0);
	a.area                                                  = 0;
	a.n_polygons = a.n_faces = a.n_points = a.n_candidates = 0;
	return a;
}

// =================================================================================================
// K3 broadphase
// =================================================================================================
struct BoxF {
	float lo[3], hi[3];
};
__device__ __forceinline__ bool overlap(const BoxF &q, const float *lo, const float *hi)
{
	return q.lo[0] <= hi[0] && q.hi[0] >= lo[0] && q.lo[1] <= hi[1] && q.hi[1] >= lo[1] && q.lo[2] <= hi[2] &&
	       q.hi[2] >= lo[2];
}

// QTET: query elements are tets of B (soft-soft) instead of triangles of B (soft-rigid)
template <bool QTET>
__global__ void __launch_bounds__(BLOCK) broadphase_kernel(PairDesc P, StepIO io)
{
	int warp = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int n_units = io.n_env * P.n_slices;
	if (warp >= n_units)
		return;
	int env = warp / P.n_slices, slice = warp - env * P.n_slices;
	Xform X_WA = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
	Xform X_WB = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
	// pair-level reject on bounding spheres
	{
		D3 ca = apply(X_WA, mk(P.A.bound_c[0], P.A.bound_c[1], P.A.bound_c[2]));
		D3 cb = apply(X_WB, mk(P.B.bound_c[0], P.B.bound_c[1], P.B.bound_c[2]));
		D3 d  = ca - cb;
		double rr = P.A.bound_r + P.B.bound_r + 1e-9;
		if (dot(d, d) > rr * rr) {
			if (lane == 0)
				P.slab_count[warp] = 0;
			return;
		}
	}
	Xform X_AB = invert_and_compose(X_WA, X_WB);
	uint2 *slab = P.slab + (size_t)warp * P.cap;
	int count   = 0;
	int q_begin = slice * P.slice_q, q_end = min(P.nq, q_begin + P.slice_q);
	unsigned lt_mask = (1u << lane) - 1u;
	for (int q0 = q_begin; q0 < q_end; q0 += 32) {
		int q      = q0 + lane;
		bool valid = q < q_end;
		BoxF box;
		if (valid) {
			double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
			const double *vp = QTET ? &P.B.tet_geom[q].v[0][0] : &P.B.tris[q].v[0][0];
			const int nv     = QTET ? 4 : 3;
#pragma unroll
			for (int i = 0; i < nv; ++i) {
				D3 p = apply(X_AB, ld3(vp + 3 * i));
				lo[0] = fmin(lo[0], p.x), lo[1] = fmin(lo[1], p.y), lo[2] = fmin(lo[2], p.z);
				hi[0] = fmax(hi[0], p.x), hi[1] = fmax(hi[1], p.y), hi[2] = fmax(hi[2], p.z);
			}
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				box.lo[a] = __double2float_rd(lo[a] - 1e-9);
				box.hi[a] = __double2float_ru(hi[a] + 1e-9);
			}
		}
		int stack[STACK_SIZE];
		int sp = 0;
		if (valid)
			stack[sp++] = 0;
		while (__any_sync(FULL_MASK, sp > 0)) {
			bool hitL = false, hitR = false;
			int eL = 0, eR = 0;
			if (sp > 0) {
				const BvhNode &nd = P.A.nodes[stack[--sp]];
				float4 a = reinterpret_cast<const float4 *>(&nd)[0], b = reinterpret_cast<const float4 *>(&nd)[1],
				       c = reinterpret_cast<const float4 *>(&nd)[2], d = reinterpret_cast<const float4 *>(&nd)[3];
				float llo[3] = { a.x, a.y, a.z }, lhi[3] = { a.w, b.x, b.y }, rlo[3] = { b.z, b.w, c.x },
				      rhi[3] = { c.y, c.z, c.w };
				int left = __float_as_int(d.x), right = __float_as_int(d.y);
				if (overlap(box, llo, lhi)) {
					if (left < 0)
						hitL = true, eL = ~left;
					else if (sp < STACK_SIZE)
						stack[sp++] = left;
					else
						atomicOr(io.flags + 1, 1);
				}
				if (overlap(box, rlo, rhi)) {
					if (right < 0)
						hitR = true, eR = ~right;
					else if (sp < STACK_SIZE)
						stack[sp++] = right;
					else
						atomicOr(io.flags + 1, 1);
				}
			}
			unsigned mL = __ballot_sync(FULL_MASK, hitL), mR = __ballot_sync(FULL_MASK, hitR);
			int nL = __popc(mL), nR = __popc(mR);
			if (hitL) {
				int pos = count + __popc(mL & lt_mask);
				if (pos < P.cap)
					slab[pos] = make_uint2((unsigned)q, (unsigned)eL);
			}
			if (hitR) {
				int pos = count + nL + __popc(mR & lt_mask);
				if (pos < P.cap)
					slab[pos] = make_uint2((unsigned)q, (unsigned)eR);
			}
			count += nL + nR;
		}
	}
	if (lane == 0) {
		if (count > P.cap) {
			atomicOr(io.flags, 1);
			count = P.cap;
		}
		P.slab_count[warp] = count;
	}
}

// =================================================================================================
// K4 soft-rigid narrowphase: one thread per (tet, triangle) candidate
// =================================================================================================
template <bool TRI>
__global__ void __launch_bounds__(BLOCK) narrow_tet_tri_kernel(PairDesc P, StepIO io)
{
	int warp = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int n_units = io.n_env * P.n_slices;
	if (warp >= n_units)
		return;
	int env = warp / P.n_slices;
	int cnt = P.slab_count[warp];
	Acc acc = zero_acc();
	if (cnt > 0) {
		Xform X_WS = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
		Xform X_WR = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
		Xform X_SR = invert_and_compose(X_WS, X_WR);
		WarpCtx ctx = make_ctx(P, io, env, X_WS, X_WR);
		const uint2 *slab = P.slab + (size_t)warp * P.cap;
		uint8_t *nvout    = P.slab_nverts + (size_t)warp * P.cap;
		const double kInf = __longlong_as_double(0x7ff0000000000000LL);
		for (int i0 = 0; i0 < cnt; i0 += 32) {
			int i       = i0 + lane;
			int nv      = 0;
			int tfaces  = 0;
			D3 W[MAXV], cW = mk(0, 0, 0);
			double e[MAXV], ec = 0;
			EmitInfo info{ env, P.index, 0, 0, i };
			if (i < cnt) {
				uint2 cand = slab[i];
				int tri = (int)cand.x, tet = (int)cand.y;
				info.elemA = tet, info.elemB = tri;
				acc.n_candidates += 1;
				const TetField &tf = P.A.tet_field[tet];
				const TriRec &tr   = P.B.tris[tri];
				D3 nS   = rot(X_SR.R, ld3(tr.n));
				D3 ghat = ld3(tf.ghat);
				if (dot(ghat, nS) > HCS_COS_ALPHA) {
					D3 bufA[MAXV], bufB[MAXV];
#pragma unroll
					for (int k = 0; k < 3; ++k)
						bufA[k] = apply(X_SR, ld3(tr.v[k]));
					int n = 3;
					n = clip_halfspace(bufA, n, ld3(tf.plane[0]), tf.plane[0][3], bufB);
					n = clip_halfspace(bufB, n, ld3(tf.plane[1]), tf.plane[1][3], bufA);
					n = clip_halfspace(bufA, n, ld3(tf.plane[2]), tf.plane[2][3], bufB);
					n = clip_halfspace(bufB, n, ld3(tf.plane[3]), tf.plane[3][3], bufA);
					n = remove_duplicates(bufA, n);
					if (n >= 3) {
						nv      = n;
						D3 grad = ld3(tf.grad);
						for (int k = 0; k < n; ++k)
							e[k] = dot(grad, bufA[k]) + tf.e0;
						integrate_polygon<TRI, false>(bufA, n, nS, grad, e, kInf, ctx, P, io, info, acc, tfaces, W, cW, ec);
					}
				}
				nvout[i] = (uint8_t)nv;
			}
			if (TRI && P.emit_tactile)
				emit_tactile(tfaces, W, cW, e, ec, P, io, info, lane);
		}
	}
	reduce_and_store(acc, P.partial + warp, lane);
}

// =================================================================================================
// K6 soft-soft narrowphase: one thread per (tet of A, tet of B) candidate
// =================================================================================================
template <bool TRI>
__global__ void __launch_bounds__(BLOCK) narrow_tet_tet_kernel(PairDesc P, StepIO io)
{
	int warp = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int n_units = io.n_env * P.n_slices;
	if (warp >= n_units)
		return;
	int env = warp / P.n_slices;
	int cnt = P.slab_count[warp];
	Acc acc = zero_acc();
	if (cnt > 0) {
		Xform X_WM = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
		Xform X_WN = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
		Xform X_MN = invert_and_compose(X_WM, X_WN);
		D3 p_NMo   = -rotT(X_MN.R, X_MN.p);
		WarpCtx ctx = make_ctx(P, io, env, X_WM, X_WN);
		const uint2 *slab = P.slab + (size_t)warp * P.cap;
		uint8_t *nvout    = P.slab_nverts + (size_t)warp * P.cap;
		for (int i0 = 0; i0 < cnt; i0 += 32) {
			int i      = i0 + lane;
			int nv     = 0;
			int tfaces = 0;
			D3 W[MAXV], cW = mk(0, 0, 0);
			double e[MAXV], ec = 0;
			EmitInfo info{ env, P.index, 0, 0, i };
			if (i < cnt) {
				uint2 cand = slab[i];
				int t1 = (int)cand.x, t0 = (int)cand.y;
				info.elemA = t0, info.elemB = t1;
				acc.n_candidates += 1;
				const TetField &f0 = P.A.tet_field[t0];
				const TetField &f1 = P.B.tet_field[t1];
				// CalcEquilibriumPlane
				D3 grad0 = ld3(f0.grad), grad1_N = ld3(f1.grad);
				double f0_Mo = f0.e0;
				D3 grad1_M   = rot(X_MN.R, grad1_N);
				double f1_Mo = dot(grad1_N, p_NMo) + f1.e0;
				D3 n_M       = grad0 - grad1_M;
				double mag   = sqrt(dot(n_M, n_M));
				bool ok      = mag > 0.0;
				D3 nhat      = mk(0, 0, 1);
				double pd    = 0;
				if (ok) {
					nhat    = n_M / mag;
					D3 p_MQ = -((f0_Mo - f1_Mo) / mag) * nhat;
					pd      = dot(nhat, p_MQ);
					ok      = dot(nhat, ld3(f0.ghat)) > HCS_COS_ALPHA;
				}
				if (ok) {
					D3 rev_N = rotT(X_MN.R, -nhat);
					ok       = dot(rev_N, ld3(f1.ghat)) > HCS_COS_ALPHA;
				}
				D3 bufA[MAXV], bufB[MAXV];
				int n = 0;
				if (ok) { // SliceTetrahedronWithPlane(tet0)
					const TetGeom &g0 = P.A.tet_geom[t0];
					double dist[4];
					int code = 0;
#pragma unroll
					for (int k = 0; k < 4; ++k) {
						dist[k] = dot(nhat, ld3(g0.v[k])) - pd;
						if (dist[k] > 0)
							code |= 1 << k;
					}
					for (int ed = 0; ed < 4; ++ed) {
						int edge = c_marching_tets[code][ed];
						if (edge < 0)
							break;
						int l0 = c_tet_edges[edge][0], l1 = c_tet_edges[edge][1];
						D3 a = ld3(g0.v[l0]), b = ld3(g0.v[l1]);
						double t  = dist[l0] / (dist[l0] - dist[l1]);
						bufA[n++] = a + t * (b - a);
					}
					n  = remove_duplicates(bufA, n);
					ok = n >= 3;
				}
				if (ok) { // clip by the four half spaces of tet1 expressed in M
					const TetGeom &g1 = P.B.tet_geom[t1];
					D3 pv[4];
#pragma unroll
					for (int k = 0; k < 4; ++k)
						pv[k] = apply(X_MN, ld3(g1.v[k]));
					const int F[4][3] = { { 1, 2, 3 }, { 0, 3, 2 }, { 0, 1, 3 }, { 0, 2, 1 } };
					D3 *in = bufA, *out = bufB;
#pragma unroll
					for (int k = 0; k < 4; ++k) {
						if (ok) {
							D3 A = pv[F[k][0]], B = pv[F[k][1]], C = pv[F[k][2]];
							D3 nh = normalized(cross(B - A, C - A));
							n     = clip_halfspace(in, n, nh, dot(nh, A), out);
							n     = remove_duplicates(out, n);
							ok    = n >= 3;
							D3 *tmp = in;
							in      = out;
							out     = tmp;
						}
					}
					if (ok) {
						nv = n;
						for (int k = 0; k < n; ++k)
							e[k] = dot(grad0, in[k]) + f0_Mo;
						double gN = -dot(grad1_M, nhat);
						integrate_polygon<TRI, false>(in, n, nhat, grad0, e, gN, ctx, P, io, info, acc, tfaces, W, cW, ec);
					}
				}
				nvout[i] = (uint8_t)nv;
			}
			if (TRI && P.emit_tactile)
				emit_tactile(tfaces, W, cW, e, ec, P, io, info, lane);
		}
	}
	reduce_and_store(acc, P.partial + warp, lane);
}

// =================================================================================================
// K5 soft-half-space narrowphase: one thread per tet of the soft geom (no candidate list needed)
// =================================================================================================
template <bool TRI>
__global__ void __launch_bounds__(BLOCK) narrow_tet_plane_kernel(PairDesc P, StepIO io)
{
	int warp = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int n_units = io.n_env * P.n_slices;
	if (warp >= n_units)
		return;
	int env = warp / P.n_slices, slice = warp - env * P.n_slices;
	Acc acc = zero_acc();
	Xform X_WS = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
	Xform X_WR = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
	Xform X_SR = invert_and_compose(X_WS, X_WR);
	D3 n_S     = mk(X_SR.R[2], X_SR.R[5], X_SR.R[8]);
	double pd  = dot(n_S, X_SR.p);
	D3 nhat_W  = rot(X_WS.R, n_S);
	WarpCtx ctx = make_ctx(P, io, env, X_WS, X_WR);
	const double kInf = __longlong_as_double(0x7ff0000000000000LL);
	int q_begin = slice * P.slice_q, q_end = min(P.nq, q_begin + P.slice_q);
	uint8_t *nvout = P.slab_nverts + (size_t)env * P.nq;
	for (int q0 = q_begin; q0 < q_end; q0 += 32) {
		int t      = q0 + lane;
		int tfaces = 0;
		D3 W[MAXV], cW = mk(0, 0, 0);
		double e[MAXV], ec = 0;
		EmitInfo info{ env, P.index, t, 0, t };
		if (t < q_end) {
			acc.n_candidates += 1;
			const TetGeom &g = P.A.tet_geom[t];
			double dist[4];
			int code = 0;
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				dist[k] = dot(n_S, ld3(g.v[k])) - pd;
				if (dist[k] > 0)
					code |= 1 << k;
			}
			int nv = 0;
			if (code != 0 && code != 15) {
				int4 gid4   = reinterpret_cast<const int4 *>(P.A.elems)[t];
				int gid[4]  = { gid4.x, gid4.y, gid4.z, gid4.w };
				D3 poly[4];
				for (int ed = 0; ed < 4; ++ed) {
					int edge = c_marching_tets[code][ed];
					if (edge < 0)
						break;
					int l0 = c_tet_edges[edge][0], l1 = c_tet_edges[edge][1];
					if (gid[l0] > gid[l1]) { // canonical direction: lower global vertex id first
						int tmp = l0;
						l0      = l1;
						l1      = tmp;
					}
					D3 a = ld3(g.v[l0]), b = ld3(g.v[l1]);
					double tt = dist[l0] / (dist[l0] - dist[l1]);
					D3 pc     = a + tt * (b - a);
					e[nv]     = g.e[l0] + tt * (g.e[l1] - g.e[l0]);
					poly[nv]  = apply(X_WS, pc);
					++nv;
				}
				D3 grad_W = rot(X_WS.R, ld3(P.A.tet_field[t].grad));
				integrate_polygon<TRI, true>(poly, nv, nhat_W, grad_W, e, kInf, ctx, P, io, info, acc, tfaces, W, cW, ec);
			}
			nvout[t] = (uint8_t)nv;
		}
		if (TRI && P.emit_tactile)
			emit_tactile(tfaces, W, cW, e, ec, P, io, info, lane);
	}
	reduce_and_store(acc, P.partial + warp, lane);
}

// =================================================================================================
// K7 finalize: fixed-order reductions
// =================================================================================================
__global__ void finalize_pairs_kernel(const PairDesc *pairs, StepIO io)
{
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= io.n_env * io.n_pairs)
		return;
	int env = idx / io.n_pairs, p = idx - env * io.n_pairs;
	const PairDesc &P = pairs[p];
	hcs_pair_result r;
	for (int k = 0; k < 3; ++k)
		r.F[k] = r.tau[k] = r.centroid[k] = 0;
	r.area = 0;
	r.gM = P.gM, r.gN = P.gN;
	r.n_polygons = r.n_faces = r.n_points = r.n_candidates = 0;
	if (P.kind != PAIR_NONE) {
		double ac[3] = { 0, 0, 0 };
		const SlicePartial *sp = P.partial + (size_t)env * P.n_slices;
		for (int s = 0; s < P.n_slices; ++s) {
			for (int k = 0; k < 3; ++k) {
				r.F[k] += sp[s].F[k];
				r.tau[k] += sp[s].tau[k];
				ac[k] += sp[s].ac[k];
			}
			r.area += sp[s].area;
			r.n_polygons += sp[s].n_polygons;
			r.n_faces += sp[s].n_faces;
			r.n_points += sp[s].n_points;
			r.n_candidates += sp[s].n_candidates;
		}
		for (int k = 0; k < 3; ++k) {
			r.F[k] *= P.sign;
			r.tau[k] *= P.sign;
			r.centroid[k] = r.area > 0 ? ac[k] / r.area : 0.0;
		}
	}
	io.pair_out[idx] = r;
}

__global__ void finalize_geoms_kernel(const PairDesc *pairs, StepIO io)
{
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= io.n_env * io.n_geoms)
		return;
	int env = idx / io.n_geoms, g = idx - env * io.n_geoms;
	double w[6] = { 0, 0, 0, 0, 0, 0 };
	for (int p = 0; p < io.n_pairs; ++p) {
		const hcs_pair_result &r = io.pair_out[(size_t)env * io.n_pairs + p];
		if (pairs[p].kind == PAIR_NONE)
			continue;
		if (r.gM == g)
			for (int k = 0; k < 3; ++k)
				w[k] += r.F[k], w[3 + k] += r.tau[k];
		if (r.gN == g)
			for (int k = 0; k < 3; ++k)
				w[k] -= r.F[k], w[3 + k] -= r.tau[k];
	}
	for (int k = 0; k < 6; ++k)
		io.geom_wrench[(size_t)idx * 6 + k] = w[k];
}

// =================================================================================================
// launchers
// =================================================================================================
static inline int blocks_for_warps(long n_warps) { return (int)((n_warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }

void launch_broadphase(const PairDesc &P, const StepIO &io, cudaStream_t s)
{
	long units = (long)io.n_env * P.n_slices;
	if (units == 0)
		return;
	if (P.kind == PAIR_SOFT_RIGID)
		broadphase_kernel<false><<<blocks_for_warps(units), BLOCK, 0, s>>>(P, io);
	else if (P.kind == PAIR_SOFT_SOFT)
		broadphase_kernel<true><<<blocks_for_warps(units), BLOCK, 0, s>>>(P, io);
}

void launch_narrowphase(const PairDesc &P, const StepIO &io, cudaStream_t s)
{
	long units = (long)io.n_env * P.n_slices;
	if (units == 0)
		return;
	int grid = blocks_for_warps(units);
	bool tri = io.representation == HCS_REP_TRIANGLE;
	switch (P.kind) {
		case PAIR_SOFT_RIGID:
			if (tri)
				narrow_tet_tri_kernel<true><<<grid, BLOCK, 0, s>>>(P, io);
			else
				narrow_tet_tri_kernel<false><<<grid, BLOCK, 0, s>>>(P, io);
			break;
		case PAIR_SOFT_SOFT:
			if (tri)
				narrow_tet_tet_kernel<true><<<grid, BLOCK, 0, s>>>(P, io);
			else
				narrow_tet_tet_kernel<false><<<grid, BLOCK, 0, s>>>(P, io);
			break;
		case PAIR_SOFT_PLANE:
			if (tri)
				narrow_tet_plane_kernel<true><<<grid, BLOCK, 0, s>>>(P, io);
			else
				narrow_tet_plane_kernel<false><<<grid, BLOCK, 0, s>>>(P, io);
			break;
		default:
			break;
	}
}

void launch_finalize(const PairDesc *d_pairs, const StepIO &io, cudaStream_t s)
{
	int n1 = io.n_env * io.n_pairs, n2 = io.n_env * io.n_geoms;
	if (n1 > 0)
		finalize_pairs_kernel<<<(n1 + 127) / 128, 128, 0, s>>>(d_pairs, io);
	if (n2 > 0)
		finalize_geoms_kernel<<<(n2 + 127) / 128, 128, 0, s>>>(d_pairs, io);
}

} // namespace hcs
