// Device code of the narrowphase kernels (kernels_step.cu: flat over the batch's candidate lists): polygon tiles in
// shared memory, Sutherland-Hodgman clip, duplicate removal,
// polygon quadrature + force law (mujoco_contact_surfaces_plugin.cpp:320-483), tactile triangle emission.
// Restates Drake mesh_intersection.cc / mesh_plane_intersection.cc / field_intersection.cc.
#pragma once
#include "dmath.cuh"
#include "hcs_internal.h"
#include "records.cuh"

namespace hcs {

#ifndef FULL_MASK
#define FULL_MASK 0xffffffffu
#endif

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
// environment read the same lines.  (Plain loads: with L1::no_allocate config 5, where one environment's block serves
// 67 000 candidates, loses 2 % of its narrowphase; configs 1 and 3 do not care: scripts/r02_run39.sh.)
struct CandCtx {
	const double *g;
	double dissipation, mu, sign;
	int apply, env, pair;
	__device__ __forceinline__ D3 v(int i) const { return xyz(ld4(g + i)); } // groups that start a 32-byte group
	__device__ __forceinline__ Xform xf(int r) const                               // R[9] + p[3] = three groups
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
			if (m != i) // (nothing dropped so far: the vertex already sits in its slot; the usual case)
				p.set(m, q);
			++m;
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

static __constant__ int c_tet_edges[6][2]      = { { 0, 1 }, { 1, 2 }, { 2, 0 }, { 0, 3 }, { 1, 3 }, { 2, 3 } };
static __constant__ int c_marching_tets[16][4] = { { -1, -1, -1, -1 }, { 0, 3, 2, -1 }, { 0, 1, 4, -1 }, { 4, 3, 2, 1 },
	                                        { 1, 2, 5, -1 },    { 0, 3, 5, 1 },  { 0, 2, 5, 4 },  { 3, 5, 4, -1 },
	                                        { 3, 4, 5, -1 },    { 4, 5, 2, 0 },  { 1, 5, 3, 0 },  { 1, 5, 2, -1 },
	                                        { 1, 2, 3, 4 },     { 0, 4, 1, -1 }, { 0, 2, 3, -1 }, { -1, -1, -1, -1 } };

__device__ __forceinline__ double pick4(const double *d, int i)
{
	return i == 0 ? d[0] : (i == 1 ? d[1] : (i == 2 ? d[2] : d[3]));
}

// optional per-face dump (PointCollision views for CPU sub-plugins); cold path, kept out of line
// Returns the face's slot when its vertices are wanted too (hcs_config.face_vertices), else -1.
static __device__ __noinline__ int dump_face(const StepIO &io, double sg, double dissipation, int env, int pair, D3 p, D3 n,
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
static __device__ __noinline__ void dump_face_vertices(double *dst, unsigned poly, int n, int a, int b, D3 cen, const double *g,
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
                                             const StepIO &io, int lane, int elemA, int elemB)
{ // (elemA, elemB): the elements of geom A (tree side) / geom B (query side; 0 for a half space) that produced the
  // polygon; with the pair and the fan index they form the triangle's canonical key
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
				t.key_hi = ((unsigned)c.pair << TRI_PAIR_SHIFT) | ((unsigned)elemB >> 2);
				t.key_lo = (((unsigned)elemB & 3u) << 30) | ((unsigned)elemA << 3) | (unsigned)i;
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

__device__ __forceinline__ Acc zero_acc()
{
	Acc a;
	a.F = a.tau = a.ac = mk(0, 0, 0);
	a.area                                                  = 0;
	a.n_polygons = a.n_faces = a.n_points = a.n_candidates = a.n_clipped = 0;
	return a;
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


#ifndef HCS_NP_PREFETCH
#define HCS_NP_PREFETCH 0
#endif
// Non-binding L1 prefetch of a line a later, dependent part of the candidate's work will gather.  Measured and left off
// (round 1): a prefetch costs the L1 data pipe as many wavefronts as the load it anticipates.
__device__ __forceinline__ void prefetch_l1(const void *p)
{
#if HCS_NP_PREFETCH
	asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

// ---- one (tet, triangle) candidate: mesh_intersection.cc ClipTriangleByTetrahedron + quadrature + force law -----------
// skip: bit f set = the broadphase's float filter saw all three triangle vertices inside half space f by more than its
// margin, so clipping by that plane cannot change a polygon made of convex combinations of them (signed distances are
// affine): the pass would copy its input, and is left out.  0 = clip against all four planes.
// buf0 / buf_stride: this lane's two polygon buffers inside the dynamic shared array.  cur: which buffer holds the result.
// Returns the number of tactile fan triangles (= polygon vertices, 0 without a polygon); nv: polygon vertex count.
template <bool TRI>
__device__ __forceinline__ int cand_tet_tri(const PairDesc &P, const StepIO &io, const CandCtx &ctx, int tri, int tet, int skip,
                                            unsigned buf0, unsigned buf_stride, Acc &acc, int &cur, D3 &cen, double &ec, int &nv)
{
	const double kInf  = __longlong_as_double(0x7ff0000000000000LL);
	const TetField *tf = P.A.tet_field + P.A.eoff(ctx.env) + tet;
	const TriVerts tr  = load_tri(P.B.tris + P.B.eoff(ctx.env) + tri);
	// the normal/gradient cull and the trivial rejects already ran in the broadphase
	const Xform X_SR = ctx.X_AB();
	D3 nS = rot(X_SR.R, tr.n);
	Poly{ buf0 }.set(0, apply(X_SR, tr.v0));
	Poly{ buf0 }.set(1, apply(X_SR, tr.v1));
	Poly{ buf0 }.set(2, apply(X_SR, tr.v2));
	int n = 3;
	cur   = 0;
#pragma unroll 1
	for (int k = 0; k < 4; ++k) {
		if ((skip >> k) & 1)
			continue;
		D4 pl = load_plane(tf, k);
		n = clip_halfspace(Poly{ buf0 + cur * buf_stride }, n, xyz(pl), pl.w, Poly{ buf0 + (cur ^ 1) * buf_stride });
		cur ^= 1;
	}
	n  = remove_duplicates(Poly{ buf0 + cur * buf_stride }, n);
	nv = 0;
	if (n < 3)
		return 0;
	nv        = n;
	D4 ge     = load_grad_e0(tf);
	D3 grad   = xyz(ge);
	double e0 = ge.w;
	// vertex pressures: the kPolygon quadrature only reads vertex 0's (the centroid's follows from the gradient), the
	// centroid fan of kTriangle and the tactile emission read them all
	const int n_press = TRI ? n : 1;
#pragma unroll 1
	for (int k = 0; k < n_press; ++k)
		PressTile{ buf0 + (cur ^ 1) * buf_stride }.set(k, dot(grad, Poly{ buf0 + cur * buf_stride }.get(k)) + e0);
	integrate_polygon<TRI, false>(Poly{ buf0 + cur * buf_stride }, n, nS, grad, PressTile{ buf0 + (cur ^ 1) * buf_stride }, kInf,
	                              ctx, io, tet, tri, acc, cen, ec);
	return n;
}

// ---- one tet of a soft geom cut by a rigid half space: mesh_half_space_intersection.cc / mesh_plane_intersection.cc -------
// Marching-tets slice, cut points along the canonical edge direction (lower global vertex id first), polygon built in
// the WORLD frame in the lane's first buffer (buf0), vertex pressures in the second (buf0 + buf_stride).
template <bool TRI>
__device__ __forceinline__ int cand_tet_plane(const PairDesc &P, const StepIO &io, const CandCtx &ctx, int t, unsigned buf0,
                                              unsigned buf_stride, Acc &acc, D3 &cen, double &ec, int &nv)
{
	const double kInf = __longlong_as_double(0x7ff0000000000000LL);
	const Poly poly{ buf0 };
	const PressTile e{ buf0 + buf_stride };
	const Xform X_WS = ctx.X_WA(), X_SR = ctx.X_AB();
	D3 n_S     = mk(X_SR.R[2], X_SR.R[5], X_SR.R[8]);
	double pd  = dot(n_S, X_SR.p);
	D3 nhat_W  = rot(X_WS.R, n_S);
	const TetVerts tg = load_tet_verts(P.A.tet_geom + P.A.eoff(ctx.env) + t);
	const D4 te       = load_tet_pressures(P.A.tet_geom + P.A.eoff(ctx.env) + t);
	double dist[4];
	int code = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		dist[k] = dot(n_S, tg.at(k)) - pd;
		if (dist[k] > 0)
			code |= 1 << k;
	}
	nv        = 0;
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
		if (TRI || nv == 0) // (kPolygon reads vertex 0's pressure only: cand_tet_tri)
			e.set(nv, pick(te, l0) + tt * (pick(te, l1) - pick(te, l0)));
		poly.set(nv, apply(X_WS, pc));
		++nv;
	}
	if (nv < 3)
		return 0;
	D3 grad_W = rot(X_WS.R, xyz(load_grad_e0(P.A.tet_field + P.A.eoff(ctx.env) + t)));
	integrate_polygon<TRI, true>(poly, nv, nhat_W, grad_W, e, kInf, ctx, io, t, 0, acc, cen, ec);
	return nv;
}

#ifndef HCS_TT_UNROLL
#define HCS_TT_UNROLL 0
#endif
// ---- one (tet of A, tet of B) candidate: field_intersection.cc CalcEquilibriumPlane + IntersectTetrahedra -----------------
template <bool TRI>
__device__ __forceinline__ int cand_tet_tet(const PairDesc &P, const StepIO &io, const CandCtx &ctx, int t1, int t0, unsigned buf0,
                                            unsigned buf_stride, Acc &acc, int &cur, D3 &cen, double &ec, int &nv)
{
	int tfaces = 0;
	cur        = 0;
	const Xform X_MN = ctx.X_AB();
	D3 p_NMo         = ctx.p_BAo();
	const size_t offA = P.A.eoff(ctx.env), offB = P.B.eoff(ctx.env);
	const TetField *f0 = P.A.tet_field + offA + t0, *f1 = P.B.tet_field + offB + t1;
	prefetch_l1(P.A.tet_geom + offA + t0); // sliced / clipped against further down, behind dependent branches
	prefetch_l1(P.B.tet_geom + offB + t1);
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
		const TetVerts g0 = load_tet_verts(P.A.tet_geom + offA + t0);
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
		const TetVerts g1 = load_tet_verts(P.B.tet_geom + offB + t1);
		D3 pv[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
			pv[k] = apply(X_MN, g1.at(k));
		// ONE copy of the pass, the face's vertices selected by k (HCS_TT_UNROLL=1: four copies): the kernel is 63 KB of SASS
		// against a 32 KB instruction cache (ncu: `no_instruction` stalls 1.1 per issue on config 3)
#if HCS_TT_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
		for (int k = 0; k < 4; ++k) {
			if (ok) {
				// outward faces {1,2,3},{0,3,2},{0,1,3},{0,2,1}
				const D3 A = k == 0 ? pv[1] : pv[0];
				const D3 B = k == 0 ? pv[2] : (k == 1 ? pv[3] : (k == 2 ? pv[1] : pv[2]));
				const D3 C = k == 0 ? pv[3] : (k == 1 ? pv[2] : (k == 2 ? pv[3] : pv[1]));
				D3 nh = normalized(cross(B - A, C - A));
				n     = clip_halfspace(Poly{ buf0 + cur * buf_stride }, n, nh, dot(nh, A), Poly{ buf0 + (cur ^ 1) * buf_stride });
				cur ^= 1;
				n  = remove_duplicates(Poly{ buf0 + cur * buf_stride }, n);
				ok = n >= 3;
			}
		}
	}
	nv = 0;
	if (ok) {
		nv = n;
		const int n_press = TRI ? n : 1; // (see cand_tet_tri)
#pragma unroll 1
		for (int k = 0; k < n_press; ++k)
			PressTile{ buf0 + (cur ^ 1) * buf_stride }.set(k, dot(grad0, Poly{ buf0 + cur * buf_stride }.get(k)) + f0_Mo);
		double gN = -dot(grad1_M, nhat);
		integrate_polygon<TRI, false>(Poly{ buf0 + cur * buf_stride }, n, nhat, grad0, PressTile{ buf0 + (cur ^ 1) * buf_stride },
		                              gN, ctx, io, t0, t1, acc, cen, ec);
		tfaces = n;
	}
	return tfaces;
}

} // namespace hcs
