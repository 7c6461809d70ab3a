// Narrowphase + reduction kernels of the hydroelastic contact engine (sm_100a, fp64, -fmad=false).
//
//   K4 narrow_tet_tri_kernel   one thread per (tet, triangle) candidate that survived the broadphase
//                              early-outs: Sutherland-Hodgman clip against the tet's four precomputed half
//                              spaces, duplicate removal, polygon quadrature (mujoco_contact_surfaces_plugin.
//                              cpp:320-409) and the force law (plugin.cpp:411-483).  Restates Drake
//                              mesh_intersection.cc.
//   K5 narrow_tet_plane_kernel one thread per tet the half space cuts (classified by the broadphase's plane units):
//                              marching-tets slice (App. A.5).
//   K6 narrow_tet_tet_kernel   one thread per (tet, tet) candidate: equal-pressure plane, slice + clip (A.6).
//   K7 finalize kernels        fixed-order reduction of the per-candidate contributions to per-pair wrenches and
//                              per-geom wrenches (replaces two mj_applyFT per face, plugin.cpp:477-482).
// All three narrowphase kernels are flat over the batch: warps pull 32-candidate chunks of the pair's candidate list.
//
// Polygon vertices live in a lane-interleaved shared-memory tile ([vertex][coord][lane], conflict free),
// loops over vertices / planes are deliberately NOT unrolled: the first version inlined everything into
// ~20k SASS instructions and stalled on instruction fetch (profiles/r01_notes.md).
// The polygon vertex arithmetic follows the oracle's (= restated Drake) operation order exactly; the
// quadrature uses the known unit normal instead of per-fan-triangle norms (differences ~1e-16 relative).
#include <algorithm>

#include "dmath.cuh"
#include "hcs_internal.h"
#include "records.cuh"

namespace hcs {

#define FULL_MASK 0xffffffffu
constexpr int NP_WARPS = 4;
constexpr int NP_BLOCK = 32 * NP_WARPS;

// per-warp shared-memory tile: two polygon buffers, lane-interleaved.  The vertex pressures of the finished polygon
// go into the buffer the clip no longer needs (MV doubles per lane of its 3 * MV): 43 KB per CTA instead of 50 KB,
// which is what lets a fifth CTA of the tet-triangle kernel fit into an SM's shared memory.
// MV1: the second buffer of a clip chain that ends in the first holds one vertex less (tet-triangle: 3 -> 4 -> 5 -> 6
// -> 7 vertices alternate between the buffers, so the second never holds more than 6; tet-tet: 4 -> ... -> 8, 7).
template <int MV, int MV1 = MV>
struct WarpTile {
	double xyz[MV][3][32];
	double xyz1[MV1][3][32];
};
#ifndef HCS_NP_TRI_CTAS // resident CTAs per SM the narrowphase kernels are compiled for (tuning sweeps: build.py)
#define HCS_NP_TRI_CTAS 4
#endif
#ifndef HCS_NP_TRI_WARPS // warps per CTA of the tet-triangle kernel
#define HCS_NP_TRI_WARPS 4
#endif
#ifndef HCS_NP_TET_CTAS // 4: 128 registers with ~100 B of spills, C3 narrowphase 1.026 -> 0.946 ms (profiles/r01_notes.md)
#define HCS_NP_TET_CTAS 4
#endif
#ifndef HCS_NP_PLANE_CTAS
#define HCS_NP_PLANE_CTAS 4
#endif

// Explicit shared-window accesses: through a generic pointer stored in a struct the compiler emitted
// generic LD/ST with 64-bit address arithmetic in the clip loop (profiles/r01_notes.md).
// All accesses index the one dynamic shared array directly, so the compiler emits LDS/STS and keeps its
// freedom to schedule them (inline-asm volatile accessors serialised the loop and were slower).
extern __shared__ __align__(16) double smem_d[];
// HCS_SMEM_INDEX: tile positions are carried as INDICES of doubles inside the dynamic shared array, not as byte offsets.
// With byte offsets every access went through `smem_d[a >> 3]`: the compiler cannot know that `a` is a multiple of 8,
// so each LDS/STS got its own add + `LOP3 & ~7` + add in front of it (three dependent integer instructions per access,
// nine per vertex; SASS of the clip loop, profiles/r01_notes.md).  With indices a vertex is one IMAD and three accesses
// with immediate offsets.
#ifndef HCS_SMEM_INDEX
#define HCS_SMEM_INDEX 1
#endif
#if HCS_SMEM_INDEX
constexpr unsigned SM_UNIT = 1u; // tile positions count doubles
__device__ __forceinline__ double lds_f64(unsigned a) { return smem_d[a]; }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { smem_d[a] = v; }
#else
constexpr unsigned SM_UNIT = 8u; // tile positions count bytes
__device__ __forceinline__ double lds_f64(unsigned a) { return smem_d[a >> 3]; }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { smem_d[a >> 3] = v; }
#endif
constexpr unsigned SM_ROW = 32u * SM_UNIT, SM_VERT = 96u * SM_UNIT; // one scalar of all 32 lanes; one vertex (x, y, z rows)

// x and y of a vertex in one 16-byte shared access (2 instead of 3 accesses per vertex, 12 % fewer instructions in
// the tet-triangle kernel).  Measured and off (scripts/sweep_r01j.sh): C1 narrowphase 0.0400 -> 0.0399 ms, C3 0.956 ->
// 0.935 ms, C5 18.7 -> 19.1 ms: the kernels wait on dependent fp64 results, not on issue slots.
#ifndef HCS_POLY_XY128
#define HCS_POLY_XY128 0
#endif
// view of one lane's polygon buffer: a = position of the buffer + lane; one vertex = 96 doubles (768 bytes).
// HCS_POLY_XY128: a vertex block holds the 32 lanes' (x, y) pairs (16 bytes each) and then their z (8 bytes each), so a
// vertex moves with one 128-bit and one 64-bit access; otherwise three 256-byte rows x, y, z.
struct Poly {
	unsigned a;
#if HCS_POLY_XY128
	__device__ __forceinline__ D3 get(int i) const
	{
		unsigned p      = a + SM_VERT * i;
		const double2 q = reinterpret_cast<const double2 *>(smem_d)[(p + SM_UNIT * (threadIdx.x & 31u)) / (2u * SM_UNIT)];
		return mk(q.x, q.y, lds_f64(p + 2u * SM_ROW));
	}
	__device__ __forceinline__ void set(int i, D3 v) const
	{
		unsigned p = a + SM_VERT * i;
		reinterpret_cast<double2 *>(smem_d)[(p + SM_UNIT * (threadIdx.x & 31u)) / (2u * SM_UNIT)] = make_double2(v.x, v.y);
		sts_f64(p + 2u * SM_ROW, v.z);
	}
#else
	__device__ __forceinline__ D3 get(int i) const
	{
		unsigned p = a + SM_VERT * i;
		return mk(lds_f64(p), lds_f64(p + SM_ROW), lds_f64(p + 2u * SM_ROW));
	}
	__device__ __forceinline__ void set(int i, D3 v) const
	{
		unsigned p = a + SM_VERT * i;
		sts_f64(p, v.x);
		sts_f64(p + SM_ROW, v.y);
		sts_f64(p + 2u * SM_ROW, v.z);
	}
#endif
};
struct PressTile { // vertex pressures of one lane
	unsigned a;
	__device__ __forceinline__ double get(int i) const { return lds_f64(a + SM_ROW * i); }
	__device__ __forceinline__ void set(int i, double v) const { sts_f64(a + SM_ROW * i, v); }
};
// position of p inside the dynamic shared array (in SM_UNITs)
__device__ __forceinline__ unsigned smem_addr(const void *p)
{
	return (unsigned)(reinterpret_cast<const char *>(p) - reinterpret_cast<const char *>(smem_d)) / (8u / SM_UNIT);
}

// Per (env, pair) data: poses, velocities, relative transform (PAIR_CTX_DOUBLES doubles in 32-byte groups, layout
// in hcs_internal.h), written by the broadphase.  They are read on demand with 256-bit loads instead of being held
// in registers: holding them across the clip loop cost ~80 registers per thread and capped the kernel at 3 warps
// per scheduler (profiles/r01_notes.md).  The lanes of a warp may belong to different environments; lanes of one
// environment read the same lines.
struct CandCtx {
	const double *g;
	double dissipation, mu, sign;
	int apply, env, pair;
	__device__ __forceinline__ D3 v(int i) const { return xyz(ld4(g + i)); } // groups that start a 32-byte group
	__device__ __forceinline__ Xform xf(int r) const                        // R[9] + p[3] = three groups
	{
		D4 a = ld4(g + r), b = ld4(g + r + 4), c = ld4(g + r + 8);
		Xform X;
		X.R[0] = a.x, X.R[1] = a.y, X.R[2] = a.z, X.R[3] = a.w;
		X.R[4] = b.x, X.R[5] = b.y, X.R[6] = b.z, X.R[7] = b.w;
		X.R[8] = c.x;
		X.p    = mk(c.y, c.z, c.w);
		return X;
	}
	__device__ __forceinline__ Xform X_WA() const { return xf(0); }        // soft geom A -> world
	__device__ __forceinline__ Xform X_AB() const { return xf(CTX_RAB); }  // geom B -> geom A
	__device__ __forceinline__ D3 p_BAo() const { return v(CTX_PBA); }     // origin of A in B
	__device__ __forceinline__ D3 xA() const                               // origin, angular, linear velocity (world)
	{
		D4 c = ld4(g + 8);
		return mk(c.y, c.z, c.w);
	}
	__device__ __forceinline__ D3 wA() const { return v(CTX_WA); }
	__device__ __forceinline__ D3 vA() const { return v(CTX_VA); }
	__device__ __forceinline__ D3 xB() const { return v(CTX_XB); }
	__device__ __forceinline__ D3 wB() const { return v(CTX_WB); }
	__device__ __forceinline__ D3 vB() const { return v(CTX_VB); }
};

struct Acc {
	D3 F, tau, ac;
	double area;
	int n_polygons, n_faces, n_points, n_candidates, n_clipped;
};

// context of one candidate of the flat narrowphase: the block the broadphase wrote for its environment
__device__ __forceinline__ CandCtx cand_ctx(const PairDesc &P, const StepIO &io, int env)
{
	CandCtx c;
	c.g           = P.pair_ctx + (size_t)env * PAIR_CTX_DOUBLES;
	c.dissipation = P.dissipation;
	c.mu          = P.mu;
	c.sign        = P.sign;
	c.apply       = io.apply_forces;
	c.env         = env;
	c.pair        = P.index;
	return c;
}

// passiveCallback force law (plugin.cpp:440-475) for one quadrature point; A = M (unswapped labelling)
template <class CTX>
__device__ __forceinline__ D3 face_force(D3 p, D3 n, double fn0, double k, const CTX &c)
{
	D3 vAq    = c.vA() + cross(c.wA(), p - c.xA());
	D3 vBq    = c.vB() + cross(c.wB(), p - c.xB());
	D3 vrel   = vAq - vBq;
	double vn = dot(vrel, n);
	double fn = fmax(0., 1. - c.dissipation * vn) * (fn0 - 0.001 * k * vn);
	if (!c.apply)
		return mk(0, 0, 0);
	D3 vt        = vrel - n * vn;
	double eps   = 1.0e-4 * 1.0e-2;
	eps          = eps * eps;
	double vslip = sqrt(dot(vt, vt) + eps);
	D3 that      = vt * (1.0 / vslip); // one division; quadrature tolerance is 1e-8, not bit parity
	double mu_r  = c.mu;
	double s     = vslip / 1.0e-4;
	if (s < 1)
		mu_r = c.mu * s * (2.0 - s);
	D3 fslip = -mu_r * that * fn;
	return fslip + fn * n;
}

// ---- Sutherland-Hodgman step: ClipPolygonByHalfSpace + CalcIntersection (mesh_intersection.cc) ----
// CalcIntersection(current, previous) with a = sd(current), b = sd(previous): wa = b / (b - a), wa * current + wb * previous
__device__ __forceinline__ D3 clip_crossing(D3 pc, double sc, D3 pprev, double sprev)
{
	double wa = sprev / (sprev - sc);
	double wb = 1.0 - wa;
	return wa * pc + wb * pprev;
}
// HCS_CLIP_DEFER=1: a plane cuts a convex polygon in at most two edges, but inside the vertex loop the crossing block (an
// IEEE division and a lerp) runs in every iteration in which any lane of the warp has a crossing.  The variant only
// reserves the output slot in the loop and computes the first two crossings of a lane after it (same operands, same
// operations: bit-identical vertices, parity green).  Measured and off (scripts/sweep_r01k.sh): C1 narrowphase 0.0400 ->
// 0.0419 ms, C3 0.958 -> 1.044 ms, C5 18.8 -> 19.4 ms: reloading the two vertices and recomputing their signed distances
// costs more than the divergent block did.  (The clip is 33 of the kernel's 40 us on C1; quadrature + force law 8.)
#ifndef HCS_CLIP_DEFER
#define HCS_CLIP_DEFER 0
#endif
// 2 would remove the eight register moves per iteration that rotate (previous vertex, previous distance) (round-2
// candidate read off the SASS, unmeasured; the loop's code doubles)
#ifndef HCS_CLIP_UNROLL
#define HCS_CLIP_UNROLL 1
#endif
constexpr int CLIP_UNROLL = HCS_CLIP_UNROLL; // (#pragma unroll takes a constant expression, not a macro)
__device__ __forceinline__ int clip_halfspace(Poly in, int n, D3 nh, double d, Poly out)
{
	if (n == 0)
		return 0;
	D3 pprev     = in.get(n - 1);
	double sprev = dot(nh, pprev) - d;
	int m        = 0;
#if HCS_CLIP_DEFER
	unsigned pending = 0; // per deferred crossing one byte: output slot | input vertex << 4
	int n_pending    = 0;
#endif
#pragma unroll CLIP_UNROLL
	for (int i = 0; i < n; ++i) {
		D3 pc     = in.get(i);
		double sc = dot(nh, pc) - d;
		bool cin = sc <= 0, pin = sprev <= 0;
		if (cin != pin) {
#if HCS_CLIP_DEFER
			if (n_pending < 2) {
				pending |= (unsigned)(m | (i << 4)) << (8 * n_pending);
				++n_pending;
				++m;
			} else
#endif
				out.set(m++, clip_crossing(pc, sc, pprev, sprev));
		}
		if (cin)
			out.set(m++, pc);
		pprev = pc;
		sprev = sc;
	}
#if HCS_CLIP_DEFER
#pragma unroll
	for (int c = 0; c < 2; ++c)
		if (c < n_pending) {
			const int slot = (pending >> (8 * c)) & 15, i = (pending >> (8 * c + 4)) & 15;
			const D3 pc = in.get(i), pp = in.get(i == 0 ? n - 1 : i - 1);
			out.set(slot, clip_crossing(pc, dot(nh, pc) - d, pp, dot(nh, pp) - d));
		}
#endif
	return m;
}

// RemoveDuplicateVertices: std::unique over consecutive near vertices, then last vs first
__device__ __forceinline__ int remove_duplicates(Poly p, int n)
{
	const double eps2 = 1e-14 * 1e-14;
	if (n == 0)
		return 0;
	int m   = 1;
	D3 last = p.get(0);
#pragma unroll 1
	for (int i = 1; i < n; ++i) {
		D3 q = p.get(i);
		D3 d = last - q;
		if (!(dot(d, d) < eps2)) {
			p.set(m++, q);
			last = q;
		}
	}
	if (m >= 3) {
		D3 d = p.get(0) - last;
		if (dot(d, d) < eps2)
			--m;
	}
	return m;
}

__constant__ int c_tet_edges[6][2]      = { { 0, 1 }, { 1, 2 }, { 2, 0 }, { 0, 3 }, { 1, 3 }, { 2, 3 } };
__constant__ int c_marching_tets[16][4] = { { -1, -1, -1, -1 }, { 0, 3, 2, -1 }, { 0, 1, 4, -1 }, { 4, 3, 2, 1 },
	                                        { 1, 2, 5, -1 },    { 0, 3, 5, 1 },  { 0, 2, 5, 4 },  { 3, 5, 4, -1 },
	                                        { 3, 4, 5, -1 },    { 4, 5, 2, 0 },  { 1, 5, 3, 0 },  { 1, 5, 2, -1 },
	                                        { 1, 2, 3, 4 },     { 0, 4, 1, -1 }, { 0, 2, 3, -1 }, { -1, -1, -1, -1 } };

__device__ __forceinline__ double pick4(const double *d, int i)
{
	return i == 0 ? d[0] : (i == 1 ? d[1] : (i == 2 ? d[2] : d[3]));
}

// optional per-face dump (PointCollision views for CPU sub-plugins); cold path, kept out of line
// Returns the face's slot when its vertices are wanted too (hcs_config.face_vertices), else -1.
__device__ __noinline__ int dump_face(const StepIO &io, double sg, double dissipation, int env, int pair, D3 p, D3 n,
                                      double fn0, double k, D3 f, int elemA, int elemB, int nverts, int face)
{
	int slot = atomicAdd(io.face_count, 1);
	if (slot >= io.max_faces)
		return -1;
	hcs_face &o = io.faces[slot];
	o.p[0] = p.x, o.p[1] = p.y, o.p[2] = p.z;
	o.n[0] = sg * n.x, o.n[1] = sg * n.y, o.n[2] = sg * n.z;
	o.fn0 = fn0, o.stiffness = k, o.damping = dissipation;
	o.f[0] = sg * f.x, o.f[1] = sg * f.y, o.f[2] = sg * f.z;
	o.env = env, o.pair = pair;
	o.elemM  = sg > 0 ? elemA : elemB;
	o.elemN  = sg > 0 ? elemB : elemA;
	o.nverts = nverts, o.face = face;
	return io.face_verts ? slot : -1;
}

// World vertices of a dumped face: what visualizeMeshElement walks (plugin.cpp:525-555).  kPolygon (b < 0): the
// polygon's n vertices; kTriangle: TriMeshBuilder's fan triangle (vertex a, vertex b, centroid).  g: the candidate's
// context block (polygon in A's frame) or NULL (polygon already in the world frame); reverse: the surface was
// swapped to (M, N) = (B, A), which reverses the winding (contact_surface.cc SwapMAndN -> ReverseFaceWinding).  Cold path, out of line.
__device__ __noinline__ void dump_face_vertices(double *dst, unsigned poly, int n, int a, int b, D3 cen, const double *g,
                                                bool reverse)
{
	Xform XW = Xform();
	if (g) {
		CandCtx c;
		c.g = g;
		XW  = c.X_WA();
	}
	const int nv = b < 0 ? n : 3;
#pragma unroll 1
	for (int i = 0; i < nv; ++i) {
		D3 v = b < 0 ? Poly{ poly }.get(i) : (i == 0 ? Poly{ poly }.get(a) : (i == 1 ? Poly{ poly }.get(b) : cen));
		if (g)
			v = apply(XW, v);
		// SwapMAndN keeps a polygon's first vertex and reverses the rest; a triangle gets its first two swapped
		const int j = !reverse ? i : (b < 0 ? (i == 0 ? 0 : nv - i) : (i == 2 ? 2 : 1 - i));
		double *o   = dst + 3 * j;
		o[0] = v.x, o[1] = v.y, o[2] = v.z;
	}
#pragma unroll 1
	for (int i = 3 * nv; i < HCS_FACE_VERTEX_STRIDE; ++i)
		dst[i] = 0.0;
}

// Quadrature + force accumulation of one contact polygon.
//   P[0..n): vertices in the builder frame (A's frame, or world when IDENT), right-handed normal nhat
//   (unit, into A); e (shared tile): vertex pressures; grad: sampled-field gradient (builder frame);
//   gN: -grad_N . nhat or +inf.  TRI selects kTriangle (centroid fan) vs kPolygon.
//   Returns the polygon centroid (builder frame) and its pressure for the tactile emission.
template <bool TRI, bool IDENT, class CTX>
__device__ __forceinline__ void integrate_polygon(Poly P, int n, D3 nhat, D3 grad, PressTile e, double gN,
                                                  const CTX &c, const StepIO &io, int elemA, int elemB, Acc &acc,
                                                  D3 &cen_out, double &ec_out)
{
	const double kInf = __longlong_as_double(0x7ff0000000000000LL);
	double gM         = dot(grad, nhat);
	const Xform XW    = IDENT ? Xform() : c.X_WA();
	D3 nW             = IDENT ? nhat : rot(XW.R, nhat);
	acc.n_polygons += 1;
	// polygon centroid (contact_surface_utility.cc CalcPolygonCentroid): fan about vertex 0, signed
	// areas measured along nhat
	D3 p0     = P.get(0);
	D3 p1     = P.get(1);
	double A2 = 0;
	D3 csum   = mk(0, 0, 0);
	D3 pi     = p1;
#pragma unroll 1
	for (int i = 1; i < n - 1; ++i) {
		D3 pn     = P.get(i + 1);
		double a2 = dot(cross(pi - p0, pn - p0), nhat);
		A2 += a2;
		csum = csum + a2 * ((p0 + pi) + pn);
		pi   = pn;
	}
	D3 cen;
	if (n == 3)
		cen = ((p0 + p1) + pi) / 3.0;
	else
		cen = A2 != 0.0 ? (TRI ? csum / (3.0 * A2) : csum * (1.0 / (3.0 * A2))) : p0; // TRI: the centroid becomes a
		                                                                              // tactile vertex, keep it exact
	double ec = e.get(0) + dot(grad, cen - p0);
	cen_out   = cen;
	ec_out    = ec;
	D3 cW     = IDENT ? cen : apply(XW, cen);
	// A face whose winding opposes nhat (only possible for a negatively oriented tet of a user mesh) gets
	// the flipped normal, like the mesh constructors that derive face normals from the winding.
	if (!TRI) {
		acc.n_faces += 1;
		double sg   = A2 < 0 ? -1.0 : 1.0;
		double area = 0.5 * (sg * A2);
		double gMf = sg * gM, gNf = gN == kInf ? gN : sg * gN;
		if (area > 0) {
			acc.area += area;
			acc.ac = acc.ac + area * cW;
		}
		if (area > 1.0e-14 && !(gMf < 1.0e-14 || gNf < 1.0e-14)) {
			double g   = gNf == kInf ? gMf : 1.0 / (1.0 / gMf + 1.0 / gNf);
			double fn0 = area * ec, k = area * g;
			D3 nf      = sg * nW;
			D3 f       = face_force(cW, nf, fn0, k, c);
			acc.F      = acc.F + f;
			acc.tau    = acc.tau + cross(cW, f);
			acc.n_points += 1;
			if (io.max_faces > 0) {
				int slot = dump_face(io, c.sign, c.dissipation, c.env, c.pair, cW, nf, fn0, k, f, elemA, elemB, n, 0);
				if (slot >= 0)
					dump_face_vertices(io.face_verts + (size_t)slot * HCS_FACE_VERTEX_STRIDE, P.a, n, 0, -1, cen,
					                   IDENT ? nullptr : c.g, c.sign < 0);
			}
		}
		return;
	}
	// kTriangle: TriMeshBuilder::AddPolygon — centroid vertex, pressure by the gradient, fan (prev,next,c)
	acc.n_faces += n;
	int cur   = n - 1;
	D3 a      = P.get(cur);
	D3 aW     = IDENT ? a : apply(XW, a);
	double ea = e.get(cur);
#pragma unroll 1
	for (int i = 0; i < n; ++i) {
		D3 b        = P.get(i);
		D3 bW       = IDENT ? b : apply(XW, b);
		double eb   = e.get(i);
		double a2   = dot(cross(b - a, cen - a), nhat);
		double sg   = a2 < 0 ? -1.0 : 1.0;
		double area = 0.5 * (sg * a2);
		double gMf = sg * gM, gNf = gN == kInf ? gN : sg * gN;
		D3 fc = ((aW + bW) + cW) * (1.0 / 3.0);
		if (area > 0) {
			acc.area += area;
			acc.ac = acc.ac + area * fc;
		}
		if (area > 1.0e-14 && !(gMf < 1.0e-14 || gNf < 1.0e-14)) {
			double g  = gNf == kInf ? gMf : 1.0 / (1.0 / gMf + 1.0 / gNf);
			double b3 = 1 / 3.;
			double pc = b3 * ea;
			pc += b3 * eb;
			pc += b3 * ec;
			double fn0 = area * pc, k = area * g;
			D3 nf      = sg * nW;
			D3 f       = face_force(fc, nf, fn0, k, c);
			acc.F      = acc.F + f;
			acc.tau    = acc.tau + cross(fc, f);
			acc.n_points += 1;
			if (io.max_faces > 0) {
				int slot = dump_face(io, c.sign, c.dissipation, c.env, c.pair, fc, nf, fn0, k, f, elemA, elemB, n, i);
				if (slot >= 0)
					dump_face_vertices(io.face_verts + (size_t)slot * HCS_FACE_VERTEX_STRIDE, P.a, n, i == 0 ? n - 1 : i - 1, i,
					                   cen, IDENT ? nullptr : c.g, c.sign < 0);
			}
		}
		a = b, aW = bW, ea = eb;
	}
}

// Warp-cooperative append of this lane's fan triangles to the tactile pool: exclusive scan over the lane
// counts, ONE atomicAdd per warp.  World vertices are recomputed from the shared tile.
template <bool IDENT, class CTX>
__device__ __forceinline__ void emit_tactile(int n_faces, Poly P, PressTile e, D3 cen, double ec, const CTX &c,
                                             const StepIO &io, int slice, int index, int lane, int elemA, int elemB)
{ // (slice, index): the emitting unit's slice and the candidate's index inside it (half space: 0 and the tet);
  // (elemA, elemB): the elements of geom A / geom B that produced the polygon
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
		const Xform XW = IDENT ? Xform() : c.X_WA();
		D3 cW     = IDENT ? cen : apply(XW, cen);
		int cur   = n_faces - 1;
		D3 aW     = IDENT ? P.get(cur) : apply(XW, P.get(cur));
		double ea = e.get(cur);
#pragma unroll 1
		for (int i = 0; i < n_faces; ++i, ++pos) {
			D3 bW     = IDENT ? P.get(i) : apply(XW, P.get(i));
			double eb = e.get(i);
			if (pos < io.max_tris) {
				// (prev, next, centroid); the M/N swap of ContactSurface reverses winding by swapping the
				// first two vertices
				bool fwd = c.sign > 0;
				D3 v0 = fwd ? aW : bW, v1 = fwd ? bW : aW;
				TactileTri t;
				t.v[0] = (float)v0.x, t.v[1] = (float)v0.y, t.v[2] = (float)v0.z;
				t.v[3] = (float)v1.x, t.v[4] = (float)v1.y, t.v[5] = (float)v1.z;
				t.v[6] = (float)cW.x, t.v[7] = (float)cW.y, t.v[8] = (float)cW.z;
				t.e[0] = fwd ? ea : eb, t.e[1] = fwd ? eb : ea, t.e[2] = ec;
				t.env        = c.env;
				t.pair_slice = ((unsigned)c.pair << TRI_SLICE_BITS) | (unsigned)slice;
				t.idx8       = (unsigned)index * 8u + (unsigned)i;
				io.tri_pool[pos] = t;
				if (io.tri_vd) { // taxel sensors sample the triangle in double
					double *vd = io.tri_vd + 9 * (size_t)pos;
					vd[0] = v0.x, vd[1] = v0.y, vd[2] = v0.z, vd[3] = v1.x, vd[4] = v1.y, vd[5] = v1.z;
					vd[6] = cW.x, vd[7] = cW.y, vd[8] = cW.z;
				}
				if (io.tri_elem)
					io.tri_elem[pos] = fwd ? make_uint2((unsigned)elemA, (unsigned)elemB) : make_uint2((unsigned)elemB, (unsigned)elemA);
			} else {
				atomicOr(io.flags, 2);
			}
			aW = bW, ea = eb;
		}
	}
}

__device__ __forceinline__ void store_partial(const Acc &t, SlicePartial *out)
{
	SlicePartial sp;
	sp.F[0] = t.F.x, sp.F[1] = t.F.y, sp.F[2] = t.F.z;
	sp.tau[0] = t.tau.x, sp.tau[1] = t.tau.y, sp.tau[2] = t.tau.z;
	sp.area = t.area;
	sp.ac[0] = t.ac.x, sp.ac[1] = t.ac.y, sp.ac[2] = t.ac.z;
	sp.n_polygons = t.n_polygons, sp.n_faces = t.n_faces, sp.n_points = t.n_points;
	sp.n_candidates = t.n_candidates, sp.n_clipped = t.n_clipped;
	sp.pad = 0;
	*out   = sp;
}

__device__ __forceinline__ Acc zero_acc()
{
	Acc a;
	a.F = a.tau = a.ac = mk(0, 0, 0);
	a.area                                                  = 0;
	a.n_polygons = a.n_faces = a.n_points = a.n_candidates = a.n_clipped = 0;
	return a;
}

// =================================================================================================
// K4 soft-rigid narrowphase: one thread per (tet, triangle) candidate
// =================================================================================================
//
// Flat over the candidates of the whole batch: the broadphase appended every (env, slice) unit's candidates to
// one list per pair; a warp grabs the next chunk of 32 consecutive records from a work counter and each lane
// reads its own environment's context block, so lane fill and load balance do not depend on how the candidates
// are spread over the environments (warp-per-environment left 15 of 32 lanes and 47 % of the resident warps busy,
// profiles/r01_notes.md).  Every candidate writes its own contribution (80 B), which K7 sums in an order that
// depends only on the candidate's index inside its unit: results do not depend on the other environments of the
// batch, on the grid size or on which warp processed the chunk.
__device__ __forceinline__ int next_chunk(int32_t *counter, int lane)
{
	int chunk = 0;
	if (lane == 0)
		chunk = atomicAdd(counter, 1);
	return __shfl_sync(FULL_MASK, chunk, 0);
}
// The round trip of the work-counter atomic is 7 % of the tet-triangle kernel's stall samples (the warp sits in the
// shuffle that broadcasts the result).  Asking for the NEXT chunk before the warp starts on the current one
// (HCS_EARLY_CLAIM=1) was measured and is off: holding the pending atomic across the body made every kernel slower
// (C1 broadphase 0.0503 -> 0.0550 ms, narrowphase 0.0398 -> 0.0416 ms; profiles/r01_notes.md).
#ifndef HCS_EARLY_CLAIM
#define HCS_EARLY_CLAIM 0
#endif
__device__ __forceinline__ int request_chunk(int32_t *counter, int lane)
{
#if HCS_EARLY_CLAIM
	return lane == 0 ? atomicAdd(counter, 1) : 0;
#else
	return 0;
#endif
}
__device__ __forceinline__ int granted_chunk(int32_t *counter, int requested, int lane)
{
#if HCS_EARLY_CLAIM
	return __shfl_sync(FULL_MASK, requested, 0);
#else
	return next_chunk(counter, lane);
#endif
}
// Non-binding L1 prefetch of a line a later, dependent part of the candidate's work will gather (the plane records
// inside the clip loop, the velocities in the force law): no register is held while the line travels.
// Measured and left off (HCS_NP_PREFETCH=1 builds it in): every lane prefetches another record, a prefetch costs the
// L1 data pipe as many wavefronts as the load it anticipates, and that pipe is the busiest unit of these kernels
// (48 % of peak, ncu); with the prefetches the tet-triangle kernel went from 0.0392 to 0.0412 ms on C1.
#ifndef HCS_NP_PREFETCH
#define HCS_NP_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l1(const void *p)
{
#if HCS_NP_PREFETCH
	asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

// candidates in the flat list, clamped to the contribution pool (overflow is reported, not UB)
__device__ __forceinline__ int flat_total(const PairDesc &P, const StepIO &io)
{
	int total = P.counters[0];
	if (total > P.contrib_cap) {
		if (blockIdx.x == 0 && threadIdx.x == 0)
			atomicOr(io.flags, 8);
		total = P.contrib_cap;
	}
	return total;
}

// Contributions: one 80-byte record per candidate (F, tau, area, area*centroid), written and read with five
// 16-byte accesses.  Measured on C1 / C3 (profiles/r01_notes.md): component-major over the whole pool put a
// candidate's ten values ~100 MB apart and the reduction ran at 350 GB/s on TLB misses; 32-candidate tiles
// [component][lane] were 45 % slower than plain records in the C1 finalize (units start at arbitrary offsets,
// so every unit touched parts of several tiles).
__device__ __forceinline__ void store_contrib(const PairDesc &P, int g, const Acc &acc)
{
	double2 *cp = reinterpret_cast<double2 *>(P.contrib + (size_t)g * 10);
	cp[0]       = make_double2(acc.F.x, acc.F.y);
	cp[1]       = make_double2(acc.F.z, acc.tau.x);
	cp[2]       = make_double2(acc.tau.y, acc.tau.z);
	cp[3]       = make_double2(acc.area, acc.ac.x);
	cp[4]       = make_double2(acc.ac.y, acc.ac.z);
}

template <bool TRI>
__global__ void __launch_bounds__(32 * HCS_NP_TRI_WARPS, HCS_NP_TRI_CTAS) narrow_tet_tri_kernel(PairDesc P, StepIO io)
{
	pdl_release(); // the finalize kernel may become resident while this grid drains
	pdl_wait();    // the broadphase grid has completed: flat list, counters and context blocks are visible
	const int lane = threadIdx.x & 31;
	WarpTile<7, 6> &T = reinterpret_cast<WarpTile<7, 6> *>(smem_d)[threadIdx.x >> 5];
	const unsigned buf0 = smem_addr(&T.xyz[0][0][lane]), buf_stride = (unsigned)sizeof(T.xyz) / (8u / SM_UNIT);
	const int total    = flat_total(P, io);
	const int n_chunks = (total + 31) >> 5;
	const double kInf  = __longlong_as_double(0x7ff0000000000000LL);
	int chunk = next_chunk(P.counters + 1, lane);
#pragma unroll 1
	while (chunk < n_chunks) {
		const int requested = request_chunk(P.counters + 1, lane);
		int g      = chunk * 32 + lane;
		int tfaces = 0;
		int cur    = 0;
		D3 cen     = mk(0, 0, 0);
		double ec  = 0;
		int env    = 0;
		uint4 rec  = make_uint4(0, 0, 0, 0); // (triangle, tet, unit, index inside the unit)
		if (g < total) {
			rec = P.flat[g];
			env = (int)rec.z / P.n_slices;
		}
		CandCtx ctx = cand_ctx(P, io, env);
		if (g < total) {
			int tri = (int)rec.x, tet = (int)rec.y;
			Acc acc = zero_acc();
			const TetField *tf = P.A.tet_field + tet;
			prefetch_l1(tf); // 192 bytes = two lines: the four planes of the clip loop, gradient and e0
			prefetch_l1(reinterpret_cast<const char *>(tf) + 128);
			prefetch_l1(ctx.g + 32); // third line of the context block: the velocities of the force law
			const TriVerts tr  = load_tri(P.B.tris + tri);
			// the normal/gradient cull and the trivial reject already ran in the broadphase
			const Xform X_SR = ctx.X_AB();
			D3 nS = rot(X_SR.R, tr.n);
			Poly{ buf0 }.set(0, apply(X_SR, tr.v0));
			Poly{ buf0 }.set(1, apply(X_SR, tr.v1));
			Poly{ buf0 }.set(2, apply(X_SR, tr.v2));
			int n = 3;
#pragma unroll 1
			for (int k = 0; k < 4; ++k) {
				D4 pl = load_plane(tf, k);
				n = clip_halfspace(Poly{ buf0 + cur * buf_stride }, n, xyz(pl), pl.w, Poly{ buf0 + (cur ^ 1) * buf_stride });
				cur ^= 1;
			}
			n      = remove_duplicates(Poly{ buf0 + cur * buf_stride }, n);
			int nv = 0;
			if (n >= 3) {
				nv        = n;
				D4 ge     = load_grad_e0(tf);
				D3 grad   = xyz(ge);
				double e0 = ge.w;
#pragma unroll 1
				for (int k = 0; k < n; ++k)
					PressTile{ buf0 + (cur ^ 1) * buf_stride }.set(k, dot(grad, Poly{ buf0 + cur * buf_stride }.get(k)) + e0);
				integrate_polygon<TRI, false>(Poly{ buf0 + cur * buf_stride }, n, nS, grad, PressTile{ buf0 + (cur ^ 1) * buf_stride },
				                              kInf, ctx, io, tet, tri, acc, cen, ec);
				tfaces = n;
			}
			// every candidate writes its record (zeros without a polygon): a 32-byte sector that is only partly
			// written in L2 has to be completed from DRAM when K7 reads it, and records share sectors
			store_contrib(P, g, acc);
			P.nverts[g] = (uint8_t)(nv | (acc.n_points << 4));
		}
		if (TRI && P.emit_tactile)
			emit_tactile<false>(tfaces, Poly{ buf0 + cur * buf_stride }, PressTile{ buf0 + (cur ^ 1) * buf_stride }, cen, ec, ctx, io,
			                    (int)rec.z - env * P.n_slices, (int)rec.w, lane, (int)rec.y, (int)rec.x);
		chunk = granted_chunk(P.counters + 1, requested, lane);
	}
}

// =================================================================================================
// K6 soft-soft narrowphase: one thread per (tet of A, tet of B) candidate
// =================================================================================================
template <bool TRI>
__global__ void __launch_bounds__(NP_BLOCK, HCS_NP_TET_CTAS) narrow_tet_tet_kernel(PairDesc P, StepIO io)
{
	pdl_release(); // the finalize kernel may become resident while this grid drains
	pdl_wait();    // the broadphase grid has completed: flat list, counters and context blocks are visible
	const int lane = threadIdx.x & 31;
	WarpTile<8, 7> &T = reinterpret_cast<WarpTile<8, 7> *>(smem_d)[threadIdx.x >> 5];
	const unsigned buf0 = smem_addr(&T.xyz[0][0][lane]), buf_stride = (unsigned)sizeof(T.xyz) / (8u / SM_UNIT);
	const int total    = flat_total(P, io);
	const int n_chunks = (total + 31) >> 5;
	int chunk = next_chunk(P.counters + 1, lane);
#pragma unroll 1
	while (chunk < n_chunks) { // flat over the candidates of the batch, see narrow_tet_tri_kernel
		const int requested = request_chunk(P.counters + 1, lane);
		int g      = chunk * 32 + lane;
		int tfaces = 0;
		int cur    = 0;
		D3 cen     = mk(0, 0, 0);
		double ec  = 0;
		int env    = 0;
		uint4 rec  = make_uint4(0, 0, 0, 0); // (tet of B, tet of A, unit, index inside the unit)
		if (g < total) {
			rec = P.flat[g];
			env = (int)rec.z / P.n_slices;
		}
		CandCtx ctx = cand_ctx(P, io, env);
		if (g < total) {
			int t1 = (int)rec.x, t0 = (int)rec.y;
			Acc acc = zero_acc();
			const Xform X_MN = ctx.X_AB();
			D3 p_NMo         = ctx.p_BAo();
			const TetField *f0 = P.A.tet_field + t0, *f1 = P.B.tet_field + t1;
			prefetch_l1(P.A.tet_geom + t0); // sliced / clipped against further down, behind dependent branches
			prefetch_l1(P.B.tet_geom + t1);
			prefetch_l1(ctx.g + 32);        // velocities of the force law
			// CalcEquilibriumPlane
			const D4 ge0 = load_grad_e0(f0), ge1 = load_grad_e0(f1);
			D3 grad0 = xyz(ge0), grad1_N = xyz(ge1);
			double f0_Mo = ge0.w;
			D3 grad1_M   = rot(X_MN.R, grad1_N);
			double f1_Mo = dot(grad1_N, p_NMo) + ge1.w;
			D3 n_M       = grad0 - grad1_M;
			double mag   = sqrt(dot(n_M, n_M));
			bool ok      = mag > 0.0;
			D3 nhat      = mk(0, 0, 1);
			double pd    = 0;
			if (ok) {
				nhat    = n_M / mag;
				D3 p_MQ = -((f0_Mo - f1_Mo) / mag) * nhat;
				pd      = dot(nhat, p_MQ);
				ok      = dot(nhat, load_ghat(f0)) > HCS_COS_ALPHA;
			}
			if (ok) {
				D3 rev_N = rotT(X_MN.R, -nhat);
				ok       = dot(rev_N, load_ghat(f1)) > HCS_COS_ALPHA;
			}
			int n = 0;
			if (ok) { // SliceTetrahedronWithPlane(tet0)
				const TetVerts g0 = load_tet_verts(P.A.tet_geom + t0);
				double dist[4];
				int code = 0;
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					dist[k] = dot(nhat, g0.at(k)) - pd;
					if (dist[k] > 0)
						code |= 1 << k;
				}
#pragma unroll 1
				for (int ed = 0; ed < 4; ++ed) {
					int edge = c_marching_tets[code][ed];
					if (edge < 0)
						break;
					int l0 = c_tet_edges[edge][0], l1 = c_tet_edges[edge][1];
					D3 a = g0.at(l0), b = g0.at(l1);
					double d0 = pick4(dist, l0), d1 = pick4(dist, l1);
					double t  = d0 / (d0 - d1);
					Poly{ buf0 }.set(n++, a + t * (b - a));
				}
				n  = remove_duplicates(Poly{ buf0 }, n);
				ok = n >= 3;
			}
			if (ok) { // clip by the four half spaces of tet1 expressed in M
				const TetVerts g1 = load_tet_verts(P.B.tet_geom + t1);
				D3 pv[4];
#pragma unroll
				for (int k = 0; k < 4; ++k)
					pv[k] = apply(X_MN, g1.at(k));
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					if (ok) {
						D3 A, B, C; // outward faces {1,2,3},{0,3,2},{0,1,3},{0,2,1}
						if (k == 0)
							A = pv[1], B = pv[2], C = pv[3];
						else if (k == 1)
							A = pv[0], B = pv[3], C = pv[2];
						else if (k == 2)
							A = pv[0], B = pv[1], C = pv[3];
						else
							A = pv[0], B = pv[2], C = pv[1];
						D3 nh = normalized(cross(B - A, C - A));
						n     = clip_halfspace(Poly{ buf0 + cur * buf_stride }, n, nh, dot(nh, A), Poly{ buf0 + (cur ^ 1) * buf_stride });
						cur ^= 1;
						n  = remove_duplicates(Poly{ buf0 + cur * buf_stride }, n);
						ok = n >= 3;
					}
				}
			}
			int nv = 0;
			if (ok) {
				nv = n;
#pragma unroll 1
				for (int k = 0; k < n; ++k)
					PressTile{ buf0 + (cur ^ 1) * buf_stride }.set(k, dot(grad0, Poly{ buf0 + cur * buf_stride }.get(k)) + f0_Mo);
				double gN = -dot(grad1_M, nhat);
				integrate_polygon<TRI, false>(Poly{ buf0 + cur * buf_stride }, n, nhat, grad0, PressTile{ buf0 + (cur ^ 1) * buf_stride },
				                              gN, ctx, io, t0, t1, acc, cen, ec);
				tfaces = n;
			}
			store_contrib(P, g, acc); // always: see narrow_tet_tri_kernel
			P.nverts[g] = (uint8_t)(nv | (acc.n_points << 4));
		}
		if (TRI && P.emit_tactile)
			emit_tactile<false>(tfaces, Poly{ buf0 + cur * buf_stride }, PressTile{ buf0 + (cur ^ 1) * buf_stride }, cen, ec, ctx, io,
			                    (int)rec.z - env * P.n_slices, (int)rec.w, lane, (int)rec.y, (int)rec.x);
		chunk = granted_chunk(P.counters + 1, requested, lane);
	}
}

// =================================================================================================
// K7 finalize, one CTA per environment.
//   phase 1: per (env, slice) unit of a candidate-list pair, one warp sums the candidate contributions in a
//            fixed order: lane l takes candidates l, l + 32, ... of the unit, then the xor-shuffle tree
//   phase 2: one thread per pair adds the unit partials in slice order -> hcs_pair_result
//   phase 3: one thread per geom adds its pairs' wrenches in pair order (replaces two mj_applyFT per face)
// =================================================================================================
struct Contrib {
	D3 F, tau, ac;
	double area;
};
__device__ __forceinline__ Contrib load_contrib(const PairDesc &P, int g)
{
	const double2 *cp = reinterpret_cast<const double2 *>(P.contrib + (size_t)g * 10);
	double2 a = cp[0], b = cp[1], c2 = cp[2], d = cp[3], e = cp[4];
	Contrib c;
	c.F    = mk(a.x, a.y, b.x);
	c.tau  = mk(b.y, c2.x, c2.y);
	c.area = d.x;
	c.ac   = mk(d.y, e.x, e.y);
	return c;
}
__device__ __forceinline__ void add_contrib(Acc &acc, const Contrib &c, int b, bool tri)
{
	int n = b & 15;
	if (n >= 3) { // candidates without a polygon wrote a record of zeros
		acc.n_polygons += 1;
		acc.n_faces += tri ? n : 1;
		acc.n_points += b >> 4;
		acc.F   = acc.F + c.F;
		acc.tau = acc.tau + c.tau;
		acc.area += c.area;
		acc.ac = acc.ac + c.ac;
	}
}

// xor-shuffle tree over groups of W consecutive lanes (W = 32: the whole warp); all 32 lanes must call it
template <int W>
__device__ __forceinline__ Acc group_sum(Acc acc)
{
	double d[10] = { acc.F.x, acc.F.y, acc.F.z, acc.tau.x, acc.tau.y, acc.tau.z, acc.area, acc.ac.x, acc.ac.y, acc.ac.z };
	int n[5]     = { acc.n_polygons, acc.n_faces, acc.n_points, acc.n_candidates, acc.n_clipped };
#pragma unroll
	for (int o = W / 2; o > 0; o >>= 1) {
#pragma unroll
		for (int k = 0; k < 10; ++k)
			d[k] += __shfl_xor_sync(FULL_MASK, d[k], o);
#pragma unroll
		for (int k = 0; k < 5; ++k)
			n[k] += __shfl_xor_sync(FULL_MASK, n[k], o);
	}
	Acc r;
	r.F = mk(d[0], d[1], d[2]), r.tau = mk(d[3], d[4], d[5]), r.area = d[6], r.ac = mk(d[7], d[8], d[9]);
	r.n_polygons = n[0], r.n_faces = n[1], r.n_points = n[2], r.n_candidates = n[3], r.n_clipped = n[4];
	return r;
}

#ifndef HCS_FIN_UNROLL
#define HCS_FIN_UNROLL 2
#endif
#ifndef HCS_FIN_SMALL_W // lanes per environment when the units hold few candidates (C1: 8 -> 0.0147 ms, 16 -> 0.0127 ms)
#define HCS_FIN_SMALL_W 16
#endif
// Totals of one (env, slice) unit on every lane of the group of W lanes that owns it (`sub` = lane inside the group;
// groups of a warp may own different units, `valid` = this group has one).  Every range but a unit's last holds whole
// 32-candidate chunks, so lane `sub` always sees the unit's candidates sub, sub + W, ... in increasing order; two
// candidates per lane are fetched together (vertex count and contribution in one round trip each) and added in order.
template <int W>
__device__ __forceinline__ Acc unit_sums(const PairDesc &P, const StepIO &io, int unit, int sub, bool valid = true)
{
	Acc acc   = zero_acc();
	int evals = 0, cnt = 0;
	if (valid) {
		int4 rg = P.unit_range[unit]; // {base, n, next, -}
		evals   = P.unit_evals[unit];
		if (W == 32 && rg.y == 0) { // nothing was clipped: most units of a multi-slice scene (warp-uniform exit)
			acc.n_candidates = evals;
			return acc;
		}
		const bool tri = io.representation == HCS_REP_TRIANGLE;
		while (rg.y > 0) {
			constexpr int U = HCS_FIN_UNROLL; // candidates per lane in flight; added in index order whatever U is
			for (int j = sub; j < rg.y; j += U * W) {
				int b[U];
				Contrib c[U];
#pragma unroll
				for (int u = 0; u < U; ++u) {
					int g   = rg.x + j + u * W;
					bool in = j + u * W < rg.y && g < P.contrib_cap;
					b[u]    = in ? P.nverts[g] : 0;
					c[u]    = load_contrib(P, in ? g : 0);
				}
#pragma unroll
				for (int u = 0; u < U; ++u)
					add_contrib(acc, c[u], b[u], tri);
			}
			cnt += rg.y;
			if (rg.z < 0)
				break;
			rg = P.ranges[rg.z];
		}
	}
	if (sub == 0) {
		acc.n_candidates = evals;
		acc.n_clipped    = cnt;
	}
	__syncwarp();
	return group_sum<W>(acc);
}

__device__ __forceinline__ void reduce_unit(const PairDesc &P, const StepIO &io, int unit, int lane)
{
	Acc t = unit_sums<32>(P, io, unit, lane);
	if (lane == 0)
		store_partial(t, P.partial + unit);
}

// =================================================================================================
// K5 soft-half-space narrowphase: one thread per tet the plane cuts (classified by the broadphase's plane units):
// marching-tets slice, cut points along the canonical edge direction, polygon built in the world frame
// =================================================================================================
template <bool TRI>
__global__ void __launch_bounds__(NP_BLOCK, HCS_NP_PLANE_CTAS) narrow_tet_plane_kernel(PairDesc P, StepIO io)
{
	pdl_release(); // the finalize kernel may become resident while this grid drains
	pdl_wait();    // the broadphase grid has completed: flat list, counters and context blocks are visible
	const int lane = threadIdx.x & 31;
	WarpTile<4, 2> &T = reinterpret_cast<WarpTile<4, 2> *>(smem_d)[threadIdx.x >> 5];
	Poly poly{ smem_addr(&T.xyz[0][0][lane]) };
	PressTile e{ smem_addr(&T.xyz1[0][0][lane]) }; // the slice needs one polygon buffer; 4 pressures fit 2 vertex rows
	const int total    = flat_total(P, io);
	const int n_chunks = (total + 31) >> 5;
	const double kInf  = __longlong_as_double(0x7ff0000000000000LL);
	int chunk = next_chunk(P.counters + 1, lane);
#pragma unroll 1
	while (chunk < n_chunks) { // flat over the cut tets of the batch, see narrow_tet_tri_kernel
		const int requested = request_chunk(P.counters + 1, lane);
		int g      = chunk * 32 + lane;
		int tfaces = 0;
		D3 cen     = mk(0, 0, 0);
		double ec  = 0;
		int env    = 0;
		uint4 rec  = make_uint4(0, 0, 0, 0); // (-, tet, unit, index inside the unit)
		if (g < total) {
			rec = P.flat[g];
			env = (int)rec.z / P.n_slices;
		}
		CandCtx ctx = cand_ctx(P, io, env);
		if (g < total) {
			const int t = (int)rec.y;
			Acc acc     = zero_acc();
			prefetch_l1(reinterpret_cast<const char *>(P.A.tet_field + t) + 128); // gradient, read after the slice
			prefetch_l1(ctx.g + 32);                                              // velocities of the force law
			const Xform X_WS = ctx.X_WA(), X_SR = ctx.X_AB();
			D3 n_S     = mk(X_SR.R[2], X_SR.R[5], X_SR.R[8]);
			double pd  = dot(n_S, X_SR.p);
			D3 nhat_W  = rot(X_WS.R, n_S);
			const TetVerts tg = load_tet_verts(P.A.tet_geom + t);
			const D4 te       = load_tet_pressures(P.A.tet_geom + t);
			double dist[4];
			int code = 0;
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				dist[k] = dot(n_S, tg.at(k)) - pd;
				if (dist[k] > 0)
					code |= 1 << k;
			}
			int nv    = 0;
			int4 gid4 = reinterpret_cast<const int4 *>(P.A.elems)[t];
#pragma unroll 1
			for (int ed = 0; ed < 4; ++ed) {
				int edge = c_marching_tets[code][ed];
				if (edge < 0)
					break;
				int l0 = c_tet_edges[edge][0], l1 = c_tet_edges[edge][1];
				int g0 = l0 == 0 ? gid4.x : (l0 == 1 ? gid4.y : (l0 == 2 ? gid4.z : gid4.w));
				int g1 = l1 == 0 ? gid4.x : (l1 == 1 ? gid4.y : (l1 == 2 ? gid4.z : gid4.w));
				if (g0 > g1) { // canonical direction: lower global vertex id first
					int tmp = l0;
					l0      = l1;
					l1      = tmp;
				}
				double d0 = pick4(dist, l0), d1 = pick4(dist, l1);
				D3 a = tg.at(l0), b = tg.at(l1);
				double tt = d0 / (d0 - d1);
				D3 pc     = a + tt * (b - a);
				e.set(nv, pick(te, l0) + tt * (pick(te, l1) - pick(te, l0)));
				poly.set(nv, apply(X_WS, pc));
				++nv;
			}
			if (nv >= 3) {
				D3 grad_W = rot(X_WS.R, xyz(load_grad_e0(P.A.tet_field + t)));
				integrate_polygon<TRI, true>(poly, nv, nhat_W, grad_W, e, kInf, ctx, io, t, 0, acc, cen, ec);
				tfaces = nv;
			}
			store_contrib(P, g, acc); // always: see narrow_tet_tri_kernel
			P.nverts[g] = (uint8_t)(nv | (acc.n_points << 4));
		}
		if (TRI && P.emit_tactile)
			emit_tactile<true>(tfaces, poly, e, cen, ec, ctx, io, (int)rec.z - env * P.n_slices, (int)rec.w, lane, (int)rec.y, 0);
		chunk = granted_chunk(P.counters + 1, requested, lane);
	}
}

// =================================================================================================
// K7 finalize: fixed-order reductions
// =================================================================================================
__device__ __forceinline__ void finalize_pair(const PairDesc *pairs, const StepIO &io, int env, int p)
{
	int idx = env * io.n_pairs + p;
	const PairDesc &P = pairs[p];
	hcs_pair_result r;
	for (int k = 0; k < 3; ++k)
		r.F[k] = r.tau[k] = r.centroid[k] = 0;
	r.area = 0;
	r.gM = P.gM, r.gN = P.gN;
	r.n_polygons = r.n_faces = r.n_points = r.n_candidates = r.n_clipped = r.reserved = 0;
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
			r.n_clipped += sp[s].n_clipped;
		}
		for (int k = 0; k < 3; ++k) {
			r.F[k] *= P.sign;
			r.tau[k] *= P.sign;
			r.centroid[k] = r.area > 0 ? ac[k] / r.area : 0.0;
		}
	}
	io.pair_out[idx] = r;
}

// A pair with many slices per environment (small batches against large trees, engine.cu build_pairs): the whole warp
// adds the unit partials, lane l the slices l, l + 32, ... in increasing order, then the fixed xor-shuffle tree.
__device__ __forceinline__ void finalize_pair_warp(const PairDesc *pairs, const StepIO &io, int env, int p, int lane)
{
	const PairDesc &P      = pairs[p];
	const SlicePartial *sp = P.partial + (size_t)env * P.n_slices;
	Acc acc = zero_acc();
	for (int s = lane; s < P.n_slices; s += 32) {
		const SlicePartial q = sp[s];
		acc.F   = acc.F + mk(q.F[0], q.F[1], q.F[2]);
		acc.tau = acc.tau + mk(q.tau[0], q.tau[1], q.tau[2]);
		acc.ac  = acc.ac + mk(q.ac[0], q.ac[1], q.ac[2]);
		acc.area += q.area;
		acc.n_polygons += q.n_polygons, acc.n_faces += q.n_faces, acc.n_points += q.n_points;
		acc.n_candidates += q.n_candidates, acc.n_clipped += q.n_clipped;
	}
	acc = group_sum<32>(acc);
	if (lane == 0) {
		hcs_pair_result r;
		r.F[0] = P.sign * acc.F.x, r.F[1] = P.sign * acc.F.y, r.F[2] = P.sign * acc.F.z;
		r.tau[0] = P.sign * acc.tau.x, r.tau[1] = P.sign * acc.tau.y, r.tau[2] = P.sign * acc.tau.z;
		r.area        = acc.area;
		r.centroid[0] = acc.area > 0 ? acc.ac.x / acc.area : 0.0;
		r.centroid[1] = acc.area > 0 ? acc.ac.y / acc.area : 0.0;
		r.centroid[2] = acc.area > 0 ? acc.ac.z / acc.area : 0.0;
		r.gM = P.gM, r.gN = P.gN;
		r.n_polygons = acc.n_polygons, r.n_faces = acc.n_faces, r.n_points = acc.n_points;
		r.n_candidates = acc.n_candidates, r.n_clipped = acc.n_clipped, r.reserved = 0;
		io.pair_out[env * io.n_pairs + p] = r;
	}
}
constexpr int FIN_COOP_SLICES = 32; // more slices than this: finalize_pair_warp

// The finalize kernel is the last kernel of a step without sensors: one thread mirrors the error flags into the
// caller's mapped pinned memory, so that the end-to-end path needs no copy after the kernels (every kernel that can
// raise a flag has finished: stream order).
__device__ __forceinline__ void publish_flags(const StepIO &io)
{
	if (io.flags_host && blockIdx.x == 0 && threadIdx.x == 0)
		for (int k = 0; k < 4; ++k)
			io.flags_host[k] = io.flags[k];
}

// phase 1 as its own grid (one warp per unit, blockIdx.y = pair) for scenes with many units per environment,
// where one CTA per environment would serialise them
__global__ void __launch_bounds__(NP_BLOCK) reduce_units_kernel(const PairDesc *pairs, StepIO io)
{
	const PairDesc &P = pairs[blockIdx.y];
	if (P.kind == PAIR_NONE)
		return;
	int warp = (blockIdx.x * NP_BLOCK + threadIdx.x) >> 5;
	if (warp < io.n_env * P.n_slices)
		reduce_unit(P, io, warp, threadIdx.x & 31);
}

// K7: one CTA per env (phases in the header comment above reduce_unit)
__global__ void __launch_bounds__(128) finalize_kernel(const PairDesc *pairs, StepIO io, int with_phase1)
{
	const int env = blockIdx.x, wid = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
	pdl_wait(); // chained behind the last narrowphase kernel when with_phase1 (no-op otherwise)
	if (with_phase1) {
		for (int p = 0; p < io.n_pairs; ++p) {
			const PairDesc &P = pairs[p];
			if (P.kind == PAIR_NONE)
				continue;
			for (int s = wid; s < P.n_slices; s += n_warps)
				reduce_unit(P, io, env * P.n_slices + s, lane);
		}
	}
	__syncthreads(); // unit partials (global) are visible to the whole CTA
	for (int p = wid; p < io.n_pairs; p += n_warps)
		if (pairs[p].kind != PAIR_NONE && pairs[p].n_slices > FIN_COOP_SLICES)
			finalize_pair_warp(pairs, io, env, p, lane);
	for (int p = threadIdx.x; p < io.n_pairs; p += blockDim.x)
		if (!(pairs[p].kind != PAIR_NONE && pairs[p].n_slices > FIN_COOP_SLICES))
			finalize_pair(pairs, io, env, p);
	__syncthreads();
	for (int g = threadIdx.x; g < io.n_geoms; g += blockDim.x) {
		double w[6] = { 0, 0, 0, 0, 0, 0 };
		for (int p = 0; p < io.n_pairs; ++p) {
			if (pairs[p].kind == PAIR_NONE)
				continue;
			const hcs_pair_result &r = io.pair_out[(size_t)env * io.n_pairs + p];
			if (r.gM == g)
				for (int k = 0; k < 3; ++k)
					w[k] += r.F[k], w[3 + k] += r.tau[k];
			if (r.gN == g)
				for (int k = 0; k < 3; ++k)
					w[k] -= r.F[k], w[3 + k] -= r.tau[k];
		}
		double *out = io.geom_wrench + ((size_t)env * io.n_geoms + g) * 6;
		for (int k = 0; k < 6; ++k)
			out[k] = w[k];
		if (io.geom_wrench_host) // end-to-end path: the caller's copy is written straight into mapped pinned memory
			for (int k = 0; k < 6; ++k)
				io.geom_wrench_host[((size_t)env * io.n_geoms + g) * 6 + k] = w[k];
	}
	publish_flags(io);
}

// K7 fast path for scenes with few units per environment: a group of W lanes (a whole warp, or 16 lanes when the
// units hold few candidates: 4 environments per warp) does all three phases of one environment with the sums in
// registers / shared memory: no block barrier, nothing written to global memory is read back (the CTA-per-environment
// kernel spent its time in that chain of dependent round trips, the warp-per-environment version in a 5-step shuffle
// tree over 15 values that mostly added zeros).  Slices in slice order, pairs in pair order, like finalize_pair + the
// geom loop above.
constexpr int FIN_WARPS = 4;
template <int W>
__global__ void __launch_bounds__(32 * FIN_WARPS) finalize_env_group_kernel(const PairDesc *pairs, StepIO io)
{
	extern __shared__ double fin_smem[]; // [FIN_WARPS * 32 / W][n_geoms][6] wrench accumulators
	pdl_wait(); // chained behind the last narrowphase kernel (no-op otherwise)
	constexpr int G = 32 / W;            // environments per warp
	const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane / W, sub = lane % W;
	const int env    = (blockIdx.x * FIN_WARPS + wid) * G + grp;
	const bool valid = env < io.n_env;
	double *w = fin_smem + (size_t)(wid * G + grp) * io.n_geoms * 6;
	for (int k = sub; k < io.n_geoms * 6; k += W)
		w[k] = 0;
	__syncwarp();
	for (int p = 0; p < io.n_pairs; ++p) {
		const PairDesc &P = pairs[p];
		hcs_pair_result r;
		for (int k = 0; k < 3; ++k)
			r.F[k] = r.tau[k] = r.centroid[k] = 0;
		r.area = 0;
		r.gM = P.gM, r.gN = P.gN;
		r.n_polygons = r.n_faces = r.n_points = r.n_candidates = r.n_clipped = r.reserved = 0;
		if (P.kind != PAIR_NONE) {
			double ac[3] = { 0, 0, 0 };
			for (int s = 0; s < P.n_slices; ++s) {
				const int unit = env * P.n_slices + s;
				Acc t          = unit_sums<W>(P, io, unit, sub, valid);
				r.F[0] += t.F.x, r.F[1] += t.F.y, r.F[2] += t.F.z;
				r.tau[0] += t.tau.x, r.tau[1] += t.tau.y, r.tau[2] += t.tau.z;
				ac[0] += t.ac.x, ac[1] += t.ac.y, ac[2] += t.ac.z;
				r.area += t.area;
				r.n_polygons += t.n_polygons, r.n_faces += t.n_faces, r.n_points += t.n_points;
				r.n_candidates += t.n_candidates, r.n_clipped += t.n_clipped;
			}
			for (int k = 0; k < 3; ++k) {
				r.F[k] *= P.sign;
				r.tau[k] *= P.sign;
				r.centroid[k] = r.area > 0 ? ac[k] / r.area : 0.0;
			}
		}
		if (sub == 0 && valid) {
			io.pair_out[(size_t)env * io.n_pairs + p] = r;
			if (P.kind != PAIR_NONE)
				for (int k = 0; k < 3; ++k) {
					w[6 * r.gM + k] += r.F[k], w[6 * r.gM + 3 + k] += r.tau[k];
					w[6 * r.gN + k] -= r.F[k], w[6 * r.gN + 3 + k] -= r.tau[k];
				}
		}
		__syncwarp();
	}
	if (valid) {
		double *out = io.geom_wrench + (size_t)env * io.n_geoms * 6;
		for (int k = sub; k < io.n_geoms * 6; k += W)
			out[k] = w[k];
		if (io.geom_wrench_host) // end-to-end path: the caller's copy is written straight into mapped pinned memory
			for (int k = sub; k < io.n_geoms * 6; k += W)
				io.geom_wrench_host[(size_t)env * io.n_geoms * 6 + k] = w[k];
	}
	publish_flags(io);
}

// =================================================================================================
// launchers
// =================================================================================================
template <class TILE, int WARPS, class K>
static void launch_np(K kernel, int grid, const PairDesc &P, const StepIO &io, cudaStream_t s, bool chained)
{
	// opt in to > 48 KB dynamic shared memory (idempotent and cheap; contexts may live on several devices)
	const int smem = (int)sizeof(TILE) * WARPS;
	cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if (chained)
		launch_chained(kernel, dim3(grid), dim3(32 * WARPS), (size_t)smem, s, P, io);
	else
		kernel<<<grid, 32 * WARPS, smem, s>>>(P, io);
}

void launch_narrowphase(const PairDesc &P, const StepIO &io, cudaStream_t s, bool chained)
{
	long units = (long)io.n_env * P.n_slices;
	if (units == 0)
		return;
	bool tri = io.representation == HCS_REP_TRIANGLE;
	// flat kernels: resident CTAs of every SM pull chunks from the work counter; never more CTAs than chunks
	long max_chunks = ((long)P.contrib_cap + 31) / 32;
	auto flat_grid  = [&](int ctas_per_sm, int warps = NP_WARPS) {
		return (int)std::max<long>(1, std::min<long>((long)io.n_sms * ctas_per_sm, (max_chunks + warps - 1) / warps));
	};
	switch (P.kind) {
		case PAIR_SOFT_RIGID:
			if (tri)
				launch_np<WarpTile<7, 6>, HCS_NP_TRI_WARPS>(narrow_tet_tri_kernel<true>, flat_grid(HCS_NP_TRI_CTAS, HCS_NP_TRI_WARPS), P, io, s, chained);
			else
				launch_np<WarpTile<7, 6>, HCS_NP_TRI_WARPS>(narrow_tet_tri_kernel<false>, flat_grid(HCS_NP_TRI_CTAS, HCS_NP_TRI_WARPS), P, io, s, chained);
			break;
		case PAIR_SOFT_SOFT:
			if (tri)
				launch_np<WarpTile<8, 7>, NP_WARPS>(narrow_tet_tet_kernel<true>, flat_grid(HCS_NP_TET_CTAS), P, io, s, chained);
			else
				launch_np<WarpTile<8, 7>, NP_WARPS>(narrow_tet_tet_kernel<false>, flat_grid(HCS_NP_TET_CTAS), P, io, s, chained);
			break;
		case PAIR_SOFT_PLANE:
			if (tri)
				launch_np<WarpTile<4, 2>, NP_WARPS>(narrow_tet_plane_kernel<true>, flat_grid(HCS_NP_PLANE_CTAS), P, io, s, chained);
			else
				launch_np<WarpTile<4, 2>, NP_WARPS>(narrow_tet_plane_kernel<false>, flat_grid(HCS_NP_PLANE_CTAS), P, io, s, chained);
			break;
		default:
			break;
	}
}

int launch_finalize(const PairDesc *d_pairs, const StepIO &io, int max_list_slices, int list_units_per_env,
                    bool small_units, cudaStream_t s, bool chained)
{
	if (io.n_env <= 0)
		return 0;
	if (list_units_per_env > 2) { // several (pair, slice) units per environment: spread phase 1 over the whole GPU
		long max_units = (long)io.n_env * max_list_slices;
		dim3 grid((unsigned)((max_units + NP_WARPS - 1) / NP_WARPS), (unsigned)io.n_pairs);
		reduce_units_kernel<<<grid, NP_BLOCK, 0, s>>>(d_pairs, io);
		finalize_kernel<<<io.n_env, 32, 0, s>>>(d_pairs, io, 0);
		return 2;
	}
	// one group of lanes per environment, everything in registers / shared memory: 16 lanes when the units are small
#ifdef HCS_FIN_FORCE_W32 // tuning sweeps
	small_units = false;
#endif
	const int W     = small_units ? HCS_FIN_SMALL_W : 32;
	const int per   = FIN_WARPS * 32 / W; // environments per CTA
	size_t smem     = (size_t)per * io.n_geoms * 6 * sizeof(double);
	if (smem <= 48 * 1024) {
		int grid = (io.n_env + per - 1) / per;
		if (W == HCS_FIN_SMALL_W && chained)
			launch_chained(finalize_env_group_kernel<HCS_FIN_SMALL_W>, dim3(grid), dim3(32 * FIN_WARPS), smem, s, d_pairs, io);
		else if (W == HCS_FIN_SMALL_W)
			finalize_env_group_kernel<HCS_FIN_SMALL_W><<<grid, 32 * FIN_WARPS, smem, s>>>(d_pairs, io);
		else if (chained)
			launch_chained(finalize_env_group_kernel<32>, dim3(grid), dim3(32 * FIN_WARPS), smem, s, d_pairs, io);
		else
			finalize_env_group_kernel<32><<<grid, 32 * FIN_WARPS, smem, s>>>(d_pairs, io);
		return 1;
	}
	// one warp per slice of a candidate-list pair (up to 4)
	int warps = std::max(1, std::min(4, max_list_slices));
	if (chained)
		launch_chained(finalize_kernel, dim3(io.n_env), dim3(32 * warps), (size_t)0, s, d_pairs, io, 1);
	else
		finalize_kernel<<<io.n_env, 32 * warps, 0, s>>>(d_pairs, io, 1);
	return 1;
}

} // namespace hcs
