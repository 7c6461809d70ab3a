// K3 broadphase — warp-cooperative LBVH traversal with work queues POOLED over several units (sm_100a).
//
// Replaces Bvh<Obb,.>::Collide inside the Drake queries the reference calls at
// mujoco_contact_surfaces_plugin.cpp:284-303 (candidate generation), and hoists the early-outs of
// mesh_intersection.cc / field_intersection.cc into the traversal so that the clipping kernel only sees pairs that need
// clipping:
//   * IsFaceNormalAlongPressureGradient / IsPlaneNormalAlongPressureGradient:  ghat . n > cos(5 pi / 8)  (a semantic
//     filter: decided in float only when clear, otherwise by the reference's fp64 expression);
//   * trivial rejects of pairs that clip to nothing (conservative float filter): all triangle
//     vertices outside one tet half space; all tet vertices on one side of the triangle's / the equal-pressure plane;
//     round 2, trees of >= 4096 tets only (PRISM_MIN_TREE): all tet vertices outside the plane through a triangle EDGE
//     along the triangle normal;
//   * behind HCS_BP_SKIP_MASK (measured, off): a 4-bit mask of tet planes that cannot cut the triangle, which the clip
//     then skips.
//
// One warp owns one (env, pair, query slice) unit.  Lane-per-query traversal left 3 of 32 lanes busy on the sphere-on-box
// scene (most query triangles die at the root), so the warp shares two LIFO queues in shared memory: node items (query
// slot, internal node) and leaf items (query slot, tet).  Every iteration all lanes pop one item each; children /
// survivors are appended with warp-ballot + popc prefix sums.
// Tried in round 2 and dropped (profiles/r02_notes.md): pooling the queues of several units in one warp (more lanes per
// iteration, fewer warps: C1 x 4096 broadphase 49 / 45 / 55 / 88 us for 1 / 2 / 4 / 8 units per warp: the kernel is bound
// by the latency of its dependent pop -> load -> test -> push chain at 16 warps per SM, not by lanes), sub-warp groups
// in lockstep (a group waits for the slowest of its warp), a flat sweep instead of the walk for 128-tet trees.
// Candidates leave in no particular order: nothing downstream depends on it (exact accumulators, hcs_internal.h).
#include <algorithm>
#include <type_traits>

#include "dmath.cuh"
#include "hcs_internal.h"
#include "records.cuh"

namespace hcs {

#define FULL_MASK 0xffffffffu
#ifndef HCS_BP_CTAS_PER_SM
#define HCS_BP_CTAS_PER_SM 4
#endif
constexpr int BP_CTAS_PER_SM = HCS_BP_CTAS_PER_SM;
constexpr int BP_WARPS = 4;
constexpr int BP_BLOCK = 32 * BP_WARPS;
constexpr int BP_SLOTS = 32;  // alive-query slots per warp: one batch
constexpr int NODE_Q   = 768; // LIFO of (slot, node); grows by <= 32 per iteration
constexpr int LEAF_Q   = 256; // (slot, tet); drained in batches of 32
constexpr int STAGE    = 256; // staged candidates per warp
// Prism test in the soft-rigid leaf filter (below).  Measured and OFF: it removes 20-29 % of the candidates that reach the
// clipper (C1 47.1 -> 37.5 per env, C5 93.4 k -> 66.5 k) but costs the leaf test ~35 warp instructions per drain:
// C1 broadphase +4 us / narrowphase -4 us, C5 broadphase +41 % instructions (+6 ms at 1024 envs) / narrowphase -2.3 ms
// (profiles/r02_notes.md).
#ifndef HCS_BP_PRISM_TEST
#define HCS_BP_PRISM_TEST 0
#endif
// Mask of tet planes the clip may skip (all triangle vertices strictly inside): measured and OFF, the narrowphase did
// not get faster (C1 0.0540 vs 0.0544 ms) and the leaf test pays ~16 instructions per hit for it.
#ifndef HCS_BP_SKIP_MASK
#define HCS_BP_SKIP_MASK 0
#endif
#ifndef HCS_SWEEP_MAX_TREE
#define HCS_SWEEP_MAX_TREE 32
#endif
constexpr int SWEEP_MAX_TREE = HCS_SWEEP_MAX_TREE; // trees up to this many tets are swept instead of walked

constexpr int ITEM_SHIFT = 26; // queue item = slot (6 bits) << 26 | node or tet index (26 bits)
constexpr unsigned ITEM_MASK = (1u << ITEM_SHIFT) - 1u;

// QF = floats kept per query (9: triangle vertices, 12: tet vertices), QX = extra rows (soft-soft); NQ / NS: capacity of
// the node queue / the candidate stage
template <int QF, int QX, int NQ = NODE_Q, int NS = STAGE>
struct __align__(16) WarpQueues {
	static constexpr int node_cap = NQ, stage_cap = NS;
	unsigned nodeq[NQ];
	unsigned leafq[LEAF_Q];
	float qbox[6][BP_SLOTS];
	float qpl[4][BP_SLOTS];  // soft-rigid: the query triangle's plane (unit normal, offset) in A's frame;
	                         // soft-soft: the query tet's pressure gradient in A's frame (0..2), its field at A's origin (3)
	float qvf[QF][BP_SLOTS]; // the query's vertices in A's frame, rounded to float (leaf filters)
	float qm[BP_SLOTS];      // soft-rigid: the filter's margin (4e-6 x the largest coordinate involved); soft-soft: that coordinate
	float qx[QX > 0 ? QX : 1][BP_SLOTS]; // soft-soft: the query tet's unit gradient in A's frame (0..2), L1 norm of its gradient (3)
	int qid[BP_SLOTS];
	int qenv[BP_SLOTS];  // flat traversal: the slot's environment (slots of one warp belong to several)
	int qev[BP_SLOTS];   // flat traversal: leaf hits of the slot's query (pair-evals started, per environment)
	int hist[BP_SLOTS];  // flat traversal: candidates per slot in the stage (flush_flat's counting sort)
	uint2 stage[NS];     // (query or slot, tet | skip << 28) waiting to be appended to the pair's flat list
	double xab[12];      // per-unit kernel: R_AB (row-major) + p_AB of the unit, for the exact fallbacks of the leaf tests
	double pba[4];       //                  origin of A in B (p_NMo of field_intersection.cc)
};
typedef WarpQueues<9, 0> QueuesRigid;
typedef WarpQueues<12, 4> QueuesSoft;
// flat traversal (bp_traverse_kernel): smaller queues, so that more warps fit into an SM's shared memory
constexpr int FT_NODE_Q = 512, FT_STAGE = 128;
typedef WarpQueues<9, 0, FT_NODE_Q, FT_STAGE> FlatQueuesRigid;
typedef WarpQueues<12, 4, FT_NODE_Q, FT_STAGE> FlatQueuesSoft;

extern __shared__ __align__(16) unsigned char bp_smem[];

// Box overlap of the query with one child of a node; for triangle queries (PLANE) the child box must also
// straddle the triangle's plane: a subtree entirely on one side of that plane cannot meet the triangle.
template <bool PLANE>
__device__ __forceinline__ bool child_overlap(const float *q, const float *pl, float lo0, float lo1, float lo2, float hi0,
                                              float hi1, float hi2)
{
	bool hit = q[0] <= hi0 && q[3] >= lo0 && q[1] <= hi1 && q[4] >= lo1 && q[2] <= hi2 && q[5] >= lo2;
	if (PLANE && hit) {
		float cx = 0.5f * (lo0 + hi0), cy = 0.5f * (lo1 + hi1), cz = 0.5f * (lo2 + hi2);
		float hx = 0.5f * (hi0 - lo0), hy = 0.5f * (hi1 - lo1), hz = 0.5f * (hi2 - lo2);
		float dist = pl[0] * cx + pl[1] * cy + pl[2] * cz - pl[3];
		float rad  = fabsf(pl[0]) * hx + fabsf(pl[1]) * hy + fabsf(pl[2]) * hz;
		// float rounding of the box/plane arithmetic is ~1e-7 relative; the margin is 1e-5 relative + 1e-7 m
		hit = fabsf(dist) <= rad * 1.00001f + 1e-5f * (fabsf(pl[3]) + fabsf(cx) + fabsf(cy) + fabsf(cz)) + 1e-7f;
	}
	return hit;
}

// per (env, pair) context block read by the narrowphase (layout: hcs_internal.h)
// part 0: the poses (from the caller's registers), part 1: the velocities (loaded here), part -1: both.  The flat prepare
// kernel gives the two parts to two threads of the environment: the one thread that wrote all 48 doubles behind its 12
// dependent velocity loads held its warp for 13 % of the kernel's stall samples (profiles/r02_ncu_c1_final.txt).
__device__ __forceinline__ void write_pair_ctx(const PairDesc &P, const StepIO &io, int env, const Xform &X_WA, const Xform &X_WB,
                                               const Xform &X_AB, D3 p_BAo, int part = -1)
{
	double *cb = P.pair_ctx + (size_t)env * PAIR_CTX_DOUBLES;
	if (part != 1) {
#pragma unroll
		for (int i = 0; i < 9; ++i) {
			cb[i]           = X_WA.R[i];
			cb[CTX_RAB + i] = X_AB.R[i];
		}
		cb[CTX_XA] = X_WA.p.x, cb[CTX_XA + 1] = X_WA.p.y, cb[CTX_XA + 2] = X_WA.p.z;
		cb[CTX_XB] = X_WB.p.x, cb[CTX_XB + 1] = X_WB.p.y, cb[CTX_XB + 2] = X_WB.p.z;
		cb[CTX_PAB] = X_AB.p.x, cb[CTX_PAB + 1] = X_AB.p.y, cb[CTX_PAB + 2] = X_AB.p.z;
		cb[CTX_PBA] = p_BAo.x, cb[CTX_PBA + 1] = p_BAo.y, cb[CTX_PBA + 2] = p_BAo.z;
		// the pad of every 32-byte group too: a sector that is only partly written is completed from DRAM when it is read
		cb[CTX_XB + 3] = cb[CTX_PBA + 3] = 0.0;
	}
	if (part != 0) {
		const double *velA = io.vel + ((size_t)env * io.n_geoms + P.gA) * 6;
		const double *velB = io.vel + ((size_t)env * io.n_geoms + P.gB) * 6;
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			cb[CTX_WA + i] = velA[i], cb[CTX_VA + i] = velA[3 + i];
			cb[CTX_WB + i] = velB[i], cb[CTX_VB + i] = velB[3 + i];
		}
		cb[CTX_WA + 3] = cb[CTX_VA + 3] = cb[CTX_WB + 3] = cb[CTX_VB + 3] = 0.0;
	}
}

template <class Q>
__device__ __forceinline__ Xform unit_xab(const Q &W)
{
	Xform X;
#pragma unroll
	for (int i = 0; i < 9; ++i)
		X.R[i] = W.xab[i];
	X.p = mk(W.xab[9], W.xab[10], W.xab[11]);
	return X;
}

// flat traversal: the slot's environment decides; the prepare kernel wrote the pair's context block
template <class Q>
__device__ __forceinline__ Xform slot_xab(const PairDesc &P, const Q &W, int s)
{
	const double *g = P.pair_ctx + (size_t)W.qenv[s] * PAIR_CTX_DOUBLES + CTX_RAB;
	const D4 a = ld4(g), b = ld4(g + 4), c = ld4(g + 8);
	Xform X;
	X.R[0] = a.x, X.R[1] = a.y, X.R[2] = a.z, X.R[3] = a.w, X.R[4] = b.x, X.R[5] = b.y, X.R[6] = b.z, X.R[7] = b.w, X.R[8] = c.x;
	X.p = mk(c.y, c.z, c.w);
	return X;
}

// ---- leaf tests ------------------------------------------------------------------------------------
// Soft-rigid.  Conservative float filter: the exact tests it stands for are early-outs, i.e. a pair
// they reject clips to nothing, so rejecting a SUBSET of those pairs changes no result, and a pair that slips through is
// clipped (to nothing) by the narrowphase.  One 128-byte TetLeaf32 line per leaf hit; the margin qm = 4e-6 x (largest
// coordinate involved) is > 5x the worst rounding of the float dot products and of the inputs' conversion.  NaNs compare
// false and fall through.  Only the gradient cull is a semantic filter: it is decided in float when clear by 1e-5,
// otherwise by the reference's fp64 expression.  s: slot of the query; skip: tet planes the clip may leave out.
template <bool FLAT, bool PRISM, class Q>
__device__ __forceinline__ bool leaf_test_rigid(const PairDesc &P, const Q &W, int s, int tet, int &skip)
{
	skip = 0;
	const int env_ = FLAT ? W.qenv[s] : 0; // (per-environment geometry goes through the flat kernels only)
	const TetLeaf32 *tl = P.A.tet_leaf32 + P.A.eoff(env_) + tet;
	const F8 l0 = ld8f(tl, 0), l1 = ld8f(tl, 1), l2 = ld8f(tl, 2), l3 = ld8f(tl, 3);
	const float nx = W.qpl[0][s], ny = W.qpl[1][s], nz = W.qpl[2][s], dtf = W.qpl[3][s], m = W.qm[s];
	const float cosg = fdot3(l2.a[0], l2.a[1], l2.a[2], nx, ny, nz);
	if (cosg < (float)HCS_COS_ALPHA - 1e-5f)
		return false;
	if (cosg < (float)HCS_COS_ALPHA + 1e-5f) { // undecided in float: the reference's expression in double
		const TriVerts tr = load_tri(P.B.tris + P.B.eoff(env_) + W.qid[s]);
		const Xform X_AB  = FLAT ? slot_xab(P, W, s) : unit_xab(W);
		if (!(dot(load_ghat(P.A.tet_field + P.A.eoff(env_) + tet), rot(X_AB.R, tr.n)) > HCS_COS_ALPHA))
			return false;
	}
	const float ax = W.qvf[0][s], ay = W.qvf[1][s], az = W.qvf[2][s], bx = W.qvf[3][s], by = W.qvf[4][s], bz = W.qvf[5][s],
	            cx = W.qvf[6][s], cy = W.qvf[7][s], cz = W.qvf[8][s];
	const float pl[4][4] = { { l0.a[0], l0.a[1], l0.a[2], l0.a[3] }, { l0.a[4], l0.a[5], l0.a[6], l0.a[7] },
		                     { l1.a[0], l1.a[1], l1.a[2], l1.a[3] }, { l1.a[4], l1.a[5], l1.a[6], l1.a[7] } };
	bool keep = true;
#pragma unroll
	for (int f = 0; f < 4; ++f) {
		float sa = fdot3(pl[f][0], pl[f][1], pl[f][2], ax, ay, az) - pl[f][3];
		float sb = fdot3(pl[f][0], pl[f][1], pl[f][2], bx, by, bz) - pl[f][3];
		float sc = fdot3(pl[f][0], pl[f][1], pl[f][2], cx, cy, cz) - pl[f][3];
		if (sa > m && sb > m && sc > m)
			keep = false;
#if HCS_BP_SKIP_MASK
		// The whole triangle is strictly inside this half space (by more than the float error): every polygon the clip
		// builds consists of convex combinations of its vertices, signed distances are affine, so clipping by this plane
		// would copy its input.  The narrowphase leaves that pass out (narrow.cuh cand_tet_tri).
		if (sa < -m && sb < -m && sc < -m)
			skip |= 1 << f;
#endif
	}
	if (!keep)
		return false;
	// tet vertices: l2.a[4..7], l3.a[0..7]
	const float tv[4][3] = { { l2.a[4], l2.a[5], l2.a[6] }, { l2.a[7], l3.a[0], l3.a[1] }, { l3.a[2], l3.a[3], l3.a[4] },
		                     { l3.a[5], l3.a[6], l3.a[7] } };
	float h0 = fdot3(nx, ny, nz, tv[0][0], tv[0][1], tv[0][2]) - dtf, h1 = fdot3(nx, ny, nz, tv[1][0], tv[1][1], tv[1][2]) - dtf;
	float h2 = fdot3(nx, ny, nz, tv[2][0], tv[2][1], tv[2][2]) - dtf, h3 = fdot3(nx, ny, nz, tv[3][0], tv[3][1], tv[3][2]) - dtf;
	if ((h0 > m && h1 > m && h2 > m && h3 > m) || (h0 < -m && h1 < -m && h2 < -m && h3 < -m))
		return false;
	if (PRISM || HCS_BP_PRISM_TEST)
	// Prism test (large trees, see bp_traverse_kernel; HCS_BP_PRISM_TEST: everywhere): the plane through a triangle edge e = q - p along the triangle normal n has the (unnormalised) outward
	// normal me = e x n, |me| = |e|.  All four tet vertices beyond it by more than m |e|_1 (>= m |e|: the float error of
	// the expression is < m |e| / 4) => tet and triangle are separated by that plane => the clip is empty.
	{
		const float px[3] = { ax, bx, cx }, py[3] = { ay, by, cy }, pz[3] = { az, bz, cz };
#pragma unroll
		for (int e = 0; e < 3; ++e) {
			const int e1   = e == 2 ? 0 : e + 1;
			const float ex = px[e1] - px[e], ey = py[e1] - py[e], ez = pz[e1] - pz[e];
			const float mx = ey * nz - ez * ny, my = ez * nx - ex * nz, mz = ex * ny - ey * nx;
			const float mm = m * (fabsf(ex) + fabsf(ey) + fabsf(ez));
			bool out = true;
#pragma unroll
			for (int k = 0; k < 4; ++k)
				out = out && fdot3(mx, my, mz, tv[k][0] - px[e], tv[k][1] - py[e], tv[k][2] - pz[e]) > mm;
			if (out)
				return false;
		}
	}
	return true;
}

// Soft-soft.  Float filter: equal-pressure plane n = grad0 - grad1, offset -(e0 - f1(Mo)) / |n| in float with explicit
// error bounds.  en bounds the direction error of the unit normal (the gradients cancel: relative error ~ eps (|grad0| +
// |grad1|) / |n|), epd the error of the offset; both blow up when the gradients nearly cancel, and then nothing is
// rejected here.  The two gradient culls are semantic filters: they are decided in float only when clear by en + 1e-5,
// otherwise the pair takes the exact path: CalcEquilibriumPlane + the two IsPlaneNormalAlongPressureGradient culls of
// field_intersection.cc (same arithmetic as the clip kernel), then: the equal-pressure plane must cut BOTH tets.
// Every comparison is written so that a NaN or an infinity leads to the exact path.
template <bool FLAT, class Q>
__device__ __forceinline__ bool leaf_test_soft(const PairDesc &P, const Q &W, int s, int tet)
{
	bool keep = true, need_exact = true;
	const int env_ = FLAT ? W.qenv[s] : 0;
	{
		const TetLeafSS32 *tl = P.A.tet_leafss32 + P.A.eoff(env_) + tet;
		const F8 l0 = ld8f(tl, 0), l1 = ld8f(tl, 1), l2 = ld8f(tl, 2);
		const float nx = l0.a[0] - W.qpl[0][s], ny = l0.a[1] - W.qpl[1][s], nz = l0.a[2] - W.qpl[2][s];
		const float mag2 = fdot3(nx, ny, nz, nx, ny, nz);
		if (mag2 > 1e-30f && mag2 < 1e30f) {
			const float r  = rsqrtf(mag2);
			const float hx = nx * r, hy = ny * r, hz = nz * r;
			const float f1o = W.qpl[3][s];
			const float G   = fabsf(l0.a[0]) + fabsf(l0.a[1]) + fabsf(l0.a[2]) + W.qx[3][s];
			const float en  = 1e-6f * G * r;
			const float pd  = -(l0.a[3] - f1o) * r;
			const float epd = 1e-6f * (fabsf(l0.a[3]) + fabsf(f1o)) * r + fabsf(pd) * (en + 1e-6f);
			const float c0  = fdot3(hx, hy, hz, l0.a[4], l0.a[5], l0.a[6]);
			const float c1  = -fdot3(hx, hy, hz, W.qx[0][s], W.qx[1][s], W.qx[2][s]);
			const float ec  = en + 1e-5f, cosa = (float)HCS_COS_ALPHA;
			if (c0 >= cosa + ec && c1 >= cosa + ec) { // both culls clearly pass
				need_exact    = false;
				const float m = epd + (en + 4e-6f) * W.qm[s];
				float h0 = fdot3(hx, hy, hz, l1.a[0], l1.a[1], l1.a[2]) - pd;
				float h1 = fdot3(hx, hy, hz, l1.a[3], l1.a[4], l1.a[5]) - pd;
				float h2 = fdot3(hx, hy, hz, l1.a[6], l1.a[7], l2.a[0]) - pd;
				float h3 = fdot3(hx, hy, hz, l2.a[1], l2.a[2], l2.a[3]) - pd;
				if ((h0 > m && h1 > m && h2 > m && h3 > m) || (h0 < -m && h1 < -m && h2 < -m && h3 < -m))
					keep = false;
				float g0 = fdot3(hx, hy, hz, W.qvf[0][s], W.qvf[1][s], W.qvf[2][s]) - pd;
				float g1 = fdot3(hx, hy, hz, W.qvf[3][s], W.qvf[4][s], W.qvf[5][s]) - pd;
				float g2 = fdot3(hx, hy, hz, W.qvf[6][s], W.qvf[7][s], W.qvf[8][s]) - pd;
				float g3 = fdot3(hx, hy, hz, W.qvf[9][s], W.qvf[10][s], W.qvf[11][s]) - pd;
				if ((g0 > m && g1 > m && g2 > m && g3 > m) || (g0 < -m && g1 < -m && g2 < -m && g3 < -m))
					keep = false;
			} else if (c0 < cosa - ec || c1 < cosa - ec) { // one cull clearly fails
				need_exact = false;
				keep       = false;
			}
		}
	}
	if (!need_exact)
		return keep;
	const Xform X_AB   = FLAT ? slot_xab(P, W, s) : unit_xab(W);
	const D3 p_BAo     = FLAT ? xyz(ld4(P.pair_ctx + (size_t)W.qenv[s] * PAIR_CTX_DOUBLES + CTX_PBA)) : mk(W.pba[0], W.pba[1], W.pba[2]);
	const TetField *f0 = P.A.tet_field + P.A.eoff(env_) + tet, *f1 = P.B.tet_field + P.B.eoff(env_) + W.qid[s];
	const D4 ge0 = load_grad_e0(f0), ge1 = load_grad_e0(f1);
	D3 grad0 = xyz(ge0), grad1_N = xyz(ge1);
	D3 grad1_M   = rot(X_AB.R, grad1_N);
	double f1_Mo = dot(grad1_N, p_BAo) + ge1.w;
	D3 n_M       = grad0 - grad1_M;
	double mag   = sqrt(dot(n_M, n_M));
	if (!(mag > 0.0))
		return false;
	D3 nhat   = n_M / mag;
	D3 p_MQ   = -((ge0.w - f1_Mo) / mag) * nhat;
	double pd = dot(nhat, p_MQ);
	if (!(dot(nhat, load_ghat(f0)) > HCS_COS_ALPHA))
		return false;
	if (!(dot(rotT(X_AB.R, -nhat), load_ghat(f1)) > HCS_COS_ALPHA))
		return false;
	const TetVerts tg = load_tet_verts(P.A.tet_geom + P.A.eoff(env_) + tet);
	double h0 = dot(nhat, tg.v0) - pd, h1 = dot(nhat, tg.v1) - pd, h2 = dot(nhat, tg.v2) - pd, h3 = dot(nhat, tg.v3) - pd;
	if ((h0 > 1e-12 && h1 > 1e-12 && h2 > 1e-12 && h3 > 1e-12) || (h0 < -1e-12 && h1 < -1e-12 && h2 < -1e-12 && h3 < -1e-12))
		return false;
	const TetVerts tq = load_tet_verts(P.B.tet_geom + P.B.eoff(env_) + W.qid[s]);
	double g0 = dot(nhat, apply(X_AB, tq.v0)) - pd, g1 = dot(nhat, apply(X_AB, tq.v1)) - pd,
	       g2 = dot(nhat, apply(X_AB, tq.v2)) - pd, g3 = dot(nhat, apply(X_AB, tq.v3)) - pd;
	if ((g0 > 1e-12 && g1 > 1e-12 && g2 > 1e-12 && g3 > 1e-12) || (g0 < -1e-12 && g1 < -1e-12 && g2 < -1e-12 && g3 < -1e-12))
		return false;
	return true;
}

// Append the staged candidates to the pair's flat list: one atomicAdd per flush.  The record carries its environment;
// where it lands does not matter (nothing downstream depends on candidate order).
template <class Q>
__device__ __forceinline__ void flush_stage(const PairDesc &P, const StepIO &io, Q &W, int lane, int env, int &n_stage)
{
	int base = 0;
	if (lane == 0)
		base = atomicAdd(P.counters, n_stage);
	base = __shfl_sync(FULL_MASK, base, 0);
	for (int j = lane; j < n_stage; j += 32) {
		if (base + j < P.contrib_cap) {
			const uint2 cd   = W.stage[j];
			P.flat[base + j] = make_uint4(cd.x, cd.y, (unsigned)env, 0u);
		} else {
			atomicOr(io.flags, 8);
		}
	}
	__syncwarp();
	n_stage = 0;
}

// KIND: 0 triangles of B against the tet tree of A, 1 tets of B against it, 2 the tets of A against a half space.
// Persistent warps: the resident CTAs of every SM pull (env, slice) units from a work counter, so the grid is never a
// fractional number of waves and a slow unit does not hold three finished warps' resources.
// KIND 2 (half space) stays on this kernel.  A flat one-thread-per-(env, tet) classification was measured in round 2 and
// dropped: every thread repeats the pose algebra a unit does once per 32 .. 512 tets (ncu, config 4 x 4096 envs: 17.1 M
// instead of 6.3 M warp instructions for the 512-tet ellipsoid, 68 vs 43 us; all four pairs 137 vs 104 us), with one
// append per warp or per CTA alike (profiles/r02_c4_launches*.csv).
template <int KIND, bool SWEEP>
__global__ void __launch_bounds__(BP_BLOCK, BP_CTAS_PER_SM) broadphase_kernel(PairDesc P, StepIO io)
{
	typedef typename std::conditional<KIND == 1, QueuesSoft, QueuesRigid>::type Queues;
	constexpr bool QTET = KIND == 1;
	Queues &W         = reinterpret_cast<Queues *>(bp_smem)[threadIdx.x >> 5];
	const int lane    = threadIdx.x & 31;
	const int n_units = io.n_env * P.n_slices;
	const unsigned lt_mask = (1u << lane) - 1u;
	const float4 *nodes4   = reinterpret_cast<const float4 *>(P.A.nodes);
	// largest coordinate of A's geometry in its own frame (bounding sphere), for the float filter's margin
	const float leaf_scale =
	    (float)(fmax(fmax(fabs(P.A.bound_c[0]), fabs(P.A.bound_c[1])), fabs(P.A.bound_c[2])) + P.A.bound_r);
	// Small trees are swept, not walked: every alive query against every tet box, 32 tets per pass (the 24-tet foam of the
	// Myrmex worlds: one pass per alive triangle).  Measured on the 128-tet sphere of config 1: 4 passes per alive triangle
	// lose to the ~27 node iterations of the walk (broadphase 0.068 vs 0.049 ms), hence the low limit.
	constexpr bool sweep = SWEEP;
	pdl_release(); // the pair's narrowphase may become resident while this grid drains (it waits for our completion)

	// (a warp's first unit is its own index in the grid, later ones come from the work counter: narrow_kernel)
	const int total_warps = (int)gridDim.x * BP_WARPS;
	bool first            = true;
	for (;;) {
		int unit = (int)blockIdx.x * BP_WARPS + (int)(threadIdx.x >> 5);
		if (!first) {
			if (n_units <= total_warps)
				break;
			if (lane == 0)
				unit = atomicAdd(P.counters + 2, 1);
			unit = total_warps + __shfl_sync(FULL_MASK, unit, 0);
		}
		first = false;
		if (unit >= n_units)
			break;
		const int env = unit / P.n_slices, slice = unit - env * P.n_slices;
		const Xform X_WA = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
		const Xform X_WB = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
		const Xform X_AB = invert_and_compose(X_WA, X_WB);
		const D3 p_BAo   = -rotT(X_AB.R, X_AB.p); // origin of A expressed in B (p_NMo of field_intersection.cc)
		bool active      = true;
		D3 n_S    = mk(0, 0, 1);
		double pd = 0;
		if (KIND != 2) { // pair-level reject on bounding spheres
			D3 ca = apply(X_WA, mk(P.A.bound_c[0], P.A.bound_c[1], P.A.bound_c[2]));
			D3 cb = apply(X_WB, mk(P.B.bound_c[0], P.B.bound_c[1], P.B.bound_c[2]));
			D3 d  = ca - cb;
			double rr = P.A.bound_r + P.B.bound_r + 1e-9;
			active    = !(dot(d, d) > rr * rr);
		} else {
			// the half space in A's frame: normal = z column of R_AB, through p_AB (mesh_half_space_intersection.cc)
			n_S = mk(X_AB.R[2], X_AB.R[5], X_AB.R[8]);
			pd  = dot(n_S, X_AB.p);
			// the whole geom above the plane: nothing can be cut
			const GeomBounds bA = geom_bounds(P.A, env);
			active = !(dot(n_S, mk(bA.c[0], bA.c[1], bA.c[2])) - pd > bA.r + 1e-9);
		}
		if (lane == 0) { // what the exact fallbacks of the leaf tests read back (rare paths)
#pragma unroll
			for (int i = 0; i < 9; ++i)
				W.xab[i] = X_AB.R[i];
			W.xab[9] = X_AB.p.x, W.xab[10] = X_AB.p.y, W.xab[11] = X_AB.p.z;
			W.pba[0] = p_BAo.x, W.pba[1] = p_BAo.y, W.pba[2] = p_BAo.z;
			if (slice == 0)
				write_pair_ctx(P, io, env, X_WA, X_WB, X_AB, p_BAo);
		}
		__syncwarp();
		const int q_begin = slice * P.slice_q, q_end = active ? min(P.nq, q_begin + P.slice_q) : q_begin;
		int n_stage = 0, n_node = 0, n_leaf = 0, evals = 0; // warp-uniform
		if (KIND == 2)
			evals = q_end - q_begin; // half space: tets examined

#pragma unroll 1
		for (int q0 = q_begin; q0 < q_end; q0 += 32) {
			const int q = q0 + lane;
			bool alive  = q < q_end;
			if (KIND == 2) {
				// ---- half space: classify tet q of A; the tets the plane cuts are the unit's candidates ----
				bool keep = false;
				if (alive) {
					const TetVerts tg = load_tet_verts(P.A.tet_geom + P.A.eoff(env) + q);
					int code = 0;
#pragma unroll
					for (int k = 0; k < 4; ++k)
						if (dot(n_S, tg.at(k)) - pd > 0)
							code |= 1 << k;
					keep = code != 0 && code != 15;
				}
				const unsigned mk_ = __ballot_sync(FULL_MASK, keep);
				if (keep)
					W.stage[n_stage + __popc(mk_ & lt_mask)] = make_uint2(0u, (unsigned)q);
				n_stage += __popc(mk_);
				__syncwarp();
				if (n_stage > STAGE - 32)
					flush_stage(P, io, W, lane, env, n_stage);
				continue;
			}
			// ---- tree kinds: transform this batch's queries, test against the root box, give the alive ones a slot ----
			double v[12];
			float box[6];
			if (alive) {
				double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
				// 12 doubles = three 256-bit loads: the tet's vertices, or the triangle's vertices + normal
				const double *vp = QTET ? reinterpret_cast<const double *>(P.B.tet_geom + q) :
				                          reinterpret_cast<const double *>(P.B.tris + q);
				const D4 r0 = ld4(vp), r1 = ld4(vp + 4), r2 = ld4(vp + 8);
				const D3 qv[4] = { mk(r0.x, r0.y, r0.z), mk(r0.w, r1.x, r1.y), mk(r1.z, r1.w, r2.x), mk(r2.y, r2.z, r2.w) };
				constexpr int nvq = QTET ? 4 : 3;
#pragma unroll
				for (int i = 0; i < nvq; ++i) {
					D3 p = apply(X_AB, qv[i]);
					v[3 * i] = p.x, v[3 * i + 1] = p.y, v[3 * i + 2] = p.z;
					lo[0] = fmin(lo[0], p.x), lo[1] = fmin(lo[1], p.y), lo[2] = fmin(lo[2], p.z);
					hi[0] = fmax(hi[0], p.x), hi[1] = fmax(hi[1], p.y), hi[2] = fmax(hi[2], p.z);
				}
				if (!QTET) {
					const D3 nS = rot(X_AB.R, qv[3]);
					v[9] = nS.x, v[10] = nS.y, v[11] = nS.z;
				}
#pragma unroll
				for (int a = 0; a < 3; ++a) {
					box[a]     = __double2float_rd(lo[a] - 1e-9);
					box[3 + a] = __double2float_ru(hi[a] + 1e-9);
				}
				alive = box[0] <= P.A.root_hi[0] && box[3] >= P.A.root_lo[0] && box[1] <= P.A.root_hi[1] &&
				        box[4] >= P.A.root_lo[1] && box[2] <= P.A.root_hi[2] && box[5] >= P.A.root_lo[2];
			}
			const unsigned m_alive = __ballot_sync(FULL_MASK, alive);
			const int n_slots      = __popc(m_alive);
			if (n_slots == 0)
				continue;
			__syncwarp();
			if (alive) {
				const int s = __popc(m_alive & lt_mask);
#pragma unroll
				for (int k = 0; k < 6; ++k)
					W.qbox[k][s] = box[k];
				float big = leaf_scale;
				constexpr int nf = QTET ? 12 : 9;
#pragma unroll
				for (int k = 0; k < nf; ++k) {
					W.qvf[k][s] = (float)v[k];
					big         = fmaxf(big, fabsf((float)v[k]));
				}
				if (!QTET) {
					W.qpl[0][s] = (float)v[9], W.qpl[1][s] = (float)v[10], W.qpl[2][s] = (float)v[11];
					W.qpl[3][s] = (float)(v[9] * v[0] + v[10] * v[1] + v[11] * v[2]);
					W.qm[s]     = 4e-6f * big + 1e-30f;
				} else {
					// what the float filter of the soft-soft leaf test needs of the query tet, computed once per query in
					// double: gradient and unit gradient rotated into A's frame, field value at A's origin
					const TetField *f1 = P.B.tet_field + q;
					const D4 ge1       = load_grad_e0(f1);
					const D3 g1M = rot(X_AB.R, xyz(ge1)), gh1M = rot(X_AB.R, load_ghat(f1));
					W.qpl[0][s] = (float)g1M.x, W.qpl[1][s] = (float)g1M.y, W.qpl[2][s] = (float)g1M.z;
					W.qpl[3][s] = (float)(dot(xyz(ge1), p_BAo) + ge1.w);
					W.qx[0][s] = (float)gh1M.x, W.qx[1][s] = (float)gh1M.y, W.qx[2][s] = (float)gh1M.z;
					W.qx[3][s] = (float)(fabs(g1M.x) + fabs(g1M.y) + fabs(g1M.z));
					W.qm[s]    = big;
				}
				W.qid[s] = q;
				if (!sweep)
					W.nodeq[s] = (unsigned)s << ITEM_SHIFT; // (slot, root)
			}
			__syncwarp();
			n_node = sweep ? 0 : n_slots;
			int sw_s = 0, sw_t = 0; // sweep position: slot, first tet of the next pass
			const int sw_slots = sweep ? n_slots : 0;

			// ---------------- shared-queue traversal ----------------
#pragma unroll 1
			while (n_node > 0 || n_leaf > 0 || sw_s < sw_slots) {
				if (n_leaf >= 32 || (n_node == 0 && sw_s >= sw_slots)) {
					// ---- drain up to 32 leaf items: leaf test, survivors are staged ----
					const int k = min(32, n_leaf);
					bool keep   = false;
					int skip    = 0, s = 0;
					unsigned tet = 0;
					if (lane < k) {
						const unsigned raw = W.leafq[n_leaf - k + lane];
						s = (int)(raw >> ITEM_SHIFT), tet = raw & ITEM_MASK;
						if constexpr (!QTET)
							keep = leaf_test_rigid<false, false>(P, W, s, (int)tet, skip);
						else
							keep = leaf_test_soft<false>(P, W, s, (int)tet);
					}
					n_leaf -= k;
					evals += k;
					const unsigned mk_ = __ballot_sync(FULL_MASK, keep);
					if (keep)
						W.stage[n_stage + __popc(mk_ & lt_mask)] = make_uint2((unsigned)W.qid[s], tet | ((unsigned)skip << CAND_MASK_SHIFT));
					n_stage += __popc(mk_);
					__syncwarp();
					if (n_stage > STAGE - 32) // the next drain may not fit
						flush_stage(P, io, W, lane, env, n_stage);
				} else if (sw_s < sw_slots) {
					// ---- sweep pass (small trees): slot sw_s against the boxes of tets sw_t .. sw_t + 31 ----
					const int t = sw_t + lane;
					bool hit    = false;
					if (t < P.n_tree) {
						float qb[6];
#pragma unroll
						for (int a = 0; a < 6; ++a)
							qb[a] = W.qbox[a][sw_s];
						float pl[4] = { 0.f, 0.f, 0.f, 0.f };
						if (!QTET) {
#pragma unroll
							for (int a = 0; a < 4; ++a)
								pl[a] = W.qpl[a][sw_s];
						}
						const F8 tb = *reinterpret_cast<const F8 *>(P.A.tet_box32 + t);
						hit = child_overlap<!QTET>(qb, pl, tb.a[0], tb.a[1], tb.a[2], tb.a[3], tb.a[4], tb.a[5]);
					}
					const unsigned mh = __ballot_sync(FULL_MASK, hit);
					if (hit)
						W.leafq[n_leaf + __popc(mh & lt_mask)] = ((unsigned)sw_s << ITEM_SHIFT) | (unsigned)t;
					n_leaf += __popc(mh);
					sw_t += 32;
					if (sw_t >= P.n_tree)
						sw_t = 0, ++sw_s;
					__syncwarp();
				} else {
					// ---- node iteration: pop k items, push <= 2k: the queue grows by <= k.  Near the capacity fewer items
					// are popped, which turns the LIFO into a depth-first walk whose stack stays below the tree depth ----
					const int k = max(1, min(min(32, n_node), NODE_Q - n_node));
					bool pushL = false, pushR = false, leafL = false, leafR = false;
					int cl = 0, cr = 0;
					unsigned s = 0;
					if (lane < k) {
						const unsigned raw = W.nodeq[n_node - k + lane];
						s                  = raw >> ITEM_SHIFT;
						float qb[6];
#pragma unroll
						for (int a = 0; a < 6; ++a)
							qb[a] = W.qbox[a][s];
						float pl[4] = { 0.f, 0.f, 0.f, 0.f };
						if (!QTET) {
#pragma unroll
							for (int a = 0; a < 4; ++a)
								pl[a] = W.qpl[a][s];
						}
						// the 64-byte node as two 256-bit loads: every lane reads another node, an LDG costs one L1 wavefront per line it
						// touches whatever its width (records.cuh), and the node loads are most of the traversal's L1 traffic
						const F8 *nd = reinterpret_cast<const F8 *>(nodes4 + 4 * (size_t)(raw & ITEM_MASK));
						const F8 n0 = nd[0], n1 = nd[1];
						const float4 a = make_float4(n0.a[0], n0.a[1], n0.a[2], n0.a[3]), b = make_float4(n0.a[4], n0.a[5], n0.a[6], n0.a[7]);
						const float4 c = make_float4(n1.a[0], n1.a[1], n1.a[2], n1.a[3]), d = make_float4(n1.a[4], n1.a[5], n1.a[6], n1.a[7]);
						cl = __float_as_int(d.x), cr = __float_as_int(d.y);
						// node layout: llo[3] lhi[3] rlo[3] rhi[3]
						if (child_overlap<!QTET>(qb, pl, a.x, a.y, a.z, a.w, b.x, b.y)) {
							leafL = cl < 0;
							pushL = !leafL;
						}
						if (child_overlap<!QTET>(qb, pl, b.z, b.w, c.x, c.y, c.z, c.w)) {
							leafR = cr < 0;
							pushR = !leafR;
						}
					}
					n_node -= k;
					__syncwarp();
					const unsigned mL = __ballot_sync(FULL_MASK, pushL), mR = __ballot_sync(FULL_MASK, pushR);
					const unsigned lL = __ballot_sync(FULL_MASK, leafL), lR = __ballot_sync(FULL_MASK, leafR);
					const int nL = __popc(mL), nR = __popc(mR), nlL = __popc(lL), nlR = __popc(lR);
					if (n_node + nL + nR > NODE_Q) { // only with k forced to 1 on a full queue: a tree deeper than NODE_Q / 32
						if (lane == 0)
							atomicOr(io.flags + 1, 1);
						n_node = 0;
						n_leaf = 0;
						break;
					}
					if (pushL)
						W.nodeq[n_node + __popc(mL & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)cl;
					if (pushR)
						W.nodeq[n_node + nL + __popc(mR & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)cr;
					// leaf items wait in the queue for an iteration or more
					if (leafL)
						W.leafq[n_leaf + __popc(lL & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)~cl;
					if (leafR)
						W.leafq[n_leaf + nlL + __popc(lR & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)~cr;
					n_node += nL + nR;
					n_leaf += nlL + nlR;
					__syncwarp();
				}
			}
		}
		if (n_stage > 0)
			flush_stage(P, io, W, lane, env, n_stage);
		// pair-evals started for this environment: LBVH leaf hits / tets classified
		if (lane == 0 && evals > 0)
			atomicAdd(reinterpret_cast<unsigned long long *>(P.accum + (size_t)env * ACC_WORDS + ACC_NEVALS), (unsigned long long)evals);
		__syncwarp();
	}
}

// =====================================================================================================
// Flat broadphase of the tree kinds (round 2): prepare + traverse.
//
// The per-unit kernel above gives every (env, query slice) its own warp: on the small scenes of configs 1 - 4 a unit has a
// dozen alive queries at most, so 12 of 32 lanes worked per iteration, every lane repeated the unit's fp64 pose algebra,
// and the kernel sat at 128 registers / 16 warps per SM.  Since the exact accumulators (hcs_internal.h) nothing
// downstream depends on the order of candidates, so queries of different environments may share a warp:
//   bp_prepare_kernel   one thread per (env, query element): relative pose, query into A's frame, root-box test; the
//                       alive ones leave as 128-byte float records (box, plane, vertices, margin, ids) in the pair's alive
//                       list (ballot + one atomicAdd per warp).  All the fp64 work of the broadphase is here.
//   (Also measured in round 2 and dropped, profiles/r02_per_lane_walk_ab.log: one alive query per LANE, each lane walking the
//   tree with a stack of its own in local memory, the warp alternating between a node phase and a leaf phase by majority:
//   no queue bookkeeping, but a query's frontier is no longer spread over the lanes.  C1 x 4096 broadphase 0.127 ms against
//   0.046, C3 0.74 against 0.60 ms, C5 x 1024 140 against 18 ms: there a few large triangles meet 131 072-tet trees.)
//   bp_traverse_kernel  persistent warps pull batches of 8 - 16 alive records, whatever their environments, and walk the
//                       tree with the two shared queues as before: float only, all slots alive, more warps per SM.
// =====================================================================================================
constexpr int AR_BOX = 0, AR_PL = 6, AR_VF = 10, AR_M = 22, AR_QX = 23, AR_QID = 27, AR_ENV = 28;
static_assert(ALIVE_WORDS == 32, "alive records are eight float4");
#ifndef HCS_FT_CTAS_PER_SM
#define HCS_FT_CTAS_PER_SM 5
#endif
constexpr int FT_CTAS_PER_SM = HCS_FT_CTAS_PER_SM; // 5 (96-register cap): C1 bp 0.0441 vs 0.0460 ms at 6, C5 equal, C3 +2 %
constexpr int PRISM_MIN_TREE = 4096;               // trees from this size on: prism test in the leaf filter, one query per batch
constexpr int PREP_BLOCK = 128;
#ifndef HCS_PREP_CTAS // resident CTAs per SM the prepare kernel is compiled for (register cap)
#define HCS_PREP_CTAS 5
#endif

// (Every thread repeats its environment's pose algebra: 111 of its ~250 fp64 operations.  Doing it once per environment
// and block - one thread per environment, results through shared memory behind a barrier - was measured and is slower:
// C1 x 4096 broadphase 0.0456 vs 0.0444 ms, C3 0.656 vs 0.620 ms, scripts/r02_run21.sh: the kernel waits on loads, not on
// the fp64 pipe, and the barrier adds a dependent stage.)
template <bool QTET>
__global__ void __launch_bounds__(PREP_BLOCK, HCS_PREP_CTAS) bp_prepare_kernel(PairDesc P, StepIO io)
{
	pdl_release();
	const long f     = (long)blockIdx.x * PREP_BLOCK + threadIdx.x;
	const long total = (long)io.n_env * P.nq;
	const int lane   = threadIdx.x & 31;
	bool alive       = f < total;
	float rec[ALIVE_WORDS];
#pragma unroll
	for (int k = 0; k < ALIVE_WORDS; ++k)
		rec[k] = 0.f;
	if (alive) {
		const int env = (int)(f / P.nq), q = (int)(f - (long)env * P.nq);
		const Xform X_WA = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
		const Xform X_WB = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
		const Xform X_AB = invert_and_compose(X_WA, X_WB);
		const D3 p_BAo   = -rotT(X_AB.R, X_AB.p); // origin of A expressed in B (p_NMo of field_intersection.cc)
		if (q == 0)
			write_pair_ctx(P, io, env, X_WA, X_WB, X_AB, p_BAo, P.nq > 1 ? 0 : -1);
		else if (q == 1)
			write_pair_ctx(P, io, env, X_WA, X_WB, X_AB, p_BAo, 1);
		const GeomBounds bA = geom_bounds(P.A, env), bB = geom_bounds(P.B, env);
		{ // pair-level reject on bounding spheres
			D3 ca = apply(X_WA, mk(bA.c[0], bA.c[1], bA.c[2]));
			D3 cb = apply(X_WB, mk(bB.c[0], bB.c[1], bB.c[2]));
			D3 d  = ca - cb;
			double rr = bA.r + bB.r + 1e-9;
			alive     = !(dot(d, d) > rr * rr);
		}
		if (alive) {
			const float leaf_scale = (float)(fmax(fmax(fabs(bA.c[0]), fabs(bA.c[1])), fabs(bA.c[2])) + bA.r);
			double v[12];
			double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
			const double *vp = QTET ? reinterpret_cast<const double *>(P.B.tet_geom + P.B.eoff(env) + q) :
			                          reinterpret_cast<const double *>(P.B.tris + P.B.eoff(env) + q);
			const D4 r0 = ld4(vp), r1 = ld4(vp + 4), r2 = ld4(vp + 8);
			const D3 qv[4] = { mk(r0.x, r0.y, r0.z), mk(r0.w, r1.x, r1.y), mk(r1.z, r1.w, r2.x), mk(r2.y, r2.z, r2.w) };
			constexpr int nvq = QTET ? 4 : 3;
#pragma unroll
			for (int i = 0; i < nvq; ++i) {
				D3 p = apply(X_AB, qv[i]);
				v[3 * i] = p.x, v[3 * i + 1] = p.y, v[3 * i + 2] = p.z;
				lo[0] = fmin(lo[0], p.x), lo[1] = fmin(lo[1], p.y), lo[2] = fmin(lo[2], p.z);
				hi[0] = fmax(hi[0], p.x), hi[1] = fmax(hi[1], p.y), hi[2] = fmax(hi[2], p.z);
			}
			if (!QTET) {
				const D3 nS = rot(X_AB.R, qv[3]);
				v[9] = nS.x, v[10] = nS.y, v[11] = nS.z;
			}
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				rec[AR_BOX + a]     = __double2float_rd(lo[a] - 1e-9);
				rec[AR_BOX + 3 + a] = __double2float_ru(hi[a] + 1e-9);
			}
			alive = rec[0] <= bA.hi[0] && rec[3] >= bA.lo[0] && rec[1] <= bA.hi[1] && rec[4] >= bA.lo[1] && rec[2] <= bA.hi[2] &&
			        rec[5] >= bA.lo[2];
			if (alive) {
				float big = leaf_scale;
				constexpr int nf = QTET ? 12 : 9;
#pragma unroll
				for (int k = 0; k < nf; ++k) {
					rec[AR_VF + k] = (float)v[k];
					big            = fmaxf(big, fabsf((float)v[k]));
				}
				if (!QTET) {
					rec[AR_PL] = (float)v[9], rec[AR_PL + 1] = (float)v[10], rec[AR_PL + 2] = (float)v[11];
					rec[AR_PL + 3] = (float)(v[9] * v[0] + v[10] * v[1] + v[11] * v[2]);
					rec[AR_M]      = 4e-6f * big + 1e-30f;
				} else {
					// what the float filter of the soft-soft leaf test needs of the query tet, computed once per query in
					// double: gradient and unit gradient rotated into A's frame, field value at A's origin
					const TetField *f1 = P.B.tet_field + P.B.eoff(env) + q;
					const D4 ge1       = load_grad_e0(f1);
					const D3 g1M = rot(X_AB.R, xyz(ge1)), gh1M = rot(X_AB.R, load_ghat(f1));
					rec[AR_PL] = (float)g1M.x, rec[AR_PL + 1] = (float)g1M.y, rec[AR_PL + 2] = (float)g1M.z;
					rec[AR_PL + 3] = (float)(dot(xyz(ge1), p_BAo) + ge1.w);
					rec[AR_QX] = (float)gh1M.x, rec[AR_QX + 1] = (float)gh1M.y, rec[AR_QX + 2] = (float)gh1M.z;
					rec[AR_QX + 3] = (float)(fabs(g1M.x) + fabs(g1M.y) + fabs(g1M.z));
					rec[AR_M]      = big;
				}
				rec[AR_QID] = __int_as_float(q);
				rec[AR_ENV] = __int_as_float(env);
			}
		}
	}
	const unsigned m_alive = __ballot_sync(FULL_MASK, alive);
	if (m_alive == 0)
		return;
	int base = 0;
	if (lane == 0)
		base = atomicAdd(P.counters + 3, __popc(m_alive));
	base = __shfl_sync(FULL_MASK, base, 0);
	if (alive) {
		const int slot = base + __popc(m_alive & ((1u << lane) - 1u));
		if (slot < P.alive_cap) {
			float4 *dst = reinterpret_cast<float4 *>(P.alive) + (size_t)slot * (ALIVE_WORDS / 4);
#pragma unroll
			for (int k = 0; k < ALIVE_WORDS / 4; ++k)
				dst[k] = make_float4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
		} else {
			atomicOr(io.flags, 8);
		}
	}
}

// Append the staged candidates (slot, tet | skip) of the current batch to the pair's flat list, grouped by slot.  The leaf
// drains take leaf items of all slots as they come, so the stage alternates between the batch's environments; written out
// in that order a 32-candidate chunk of the narrowphase holds a dozen short runs of equal environments, and every run
// costs it 22 integer atomics (accumulate_chunk).  A counting sort by slot inside the warp (histogram with shared-memory
// atomics, one warp scan, scatter) makes the runs as long as the slots' candidate lists.  FT_STAGE <= 4 x 32 entries.
#ifndef HCS_FLUSH_SORTED
#define HCS_FLUSH_SORTED 1
#endif
// Measured on one box (scripts/r02_run26.sh): C1 x 4096 narrowphase 0.0430 -> 0.0405 ms, broadphase 0.0443 -> 0.0458 ms
// (value 44.4 -> 44.9 M); soft-soft batches (hundreds of candidates per slot: the runs are long anyway) and single-slot
// batches of large trees only pay for it (C3 broadphase 0.631 -> 0.656 ms), so they keep the staged order (`sorted`).
template <class Q>
__device__ __forceinline__ void flush_flat(const PairDesc &P, const StepIO &io, Q &W, int lane, int &n_stage, bool sorted)
{
	int base = 0;
	if (lane == 0)
		base = atomicAdd(P.counters, n_stage);
	base = __shfl_sync(FULL_MASK, base, 0);
#if HCS_FLUSH_SORTED
	static_assert(Q::stage_cap <= 4 * 32, "flush_flat keeps four staged entries per lane");
	if (!sorted) {
		for (int j = lane; j < n_stage; j += 32) {
			if (base + j < P.contrib_cap) {
				const uint2 cd   = W.stage[j];
				P.flat[base + j] = make_uint4((unsigned)W.qid[cd.x], cd.y, (unsigned)W.qenv[cd.x], 0u);
			} else {
				atomicOr(io.flags, 8);
			}
		}
		__syncwarp();
		n_stage = 0;
		return;
	}
	W.hist[lane] = 0;
	__syncwarp();
	int rank[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const int j = lane + 32 * k;
		rank[k]     = j < n_stage ? atomicAdd(&W.hist[W.stage[j].x], 1) : 0; // position inside the slot's group
	}
	__syncwarp();
	const int cnt = W.hist[lane];
	int incl      = cnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int t = __shfl_up_sync(FULL_MASK, incl, o);
		if (lane >= o)
			incl += t;
	}
	__syncwarp();
	W.hist[lane] = incl - cnt; // first position of the slot's group
	__syncwarp();
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const int j = lane + 32 * k;
		if (j < n_stage) {
			const uint2 cd = W.stage[j];
			const int dst  = base + W.hist[cd.x] + rank[k];
			if (dst < P.contrib_cap)
				P.flat[dst] = make_uint4((unsigned)W.qid[cd.x], cd.y, (unsigned)W.qenv[cd.x], 0u);
			else
				atomicOr(io.flags, 8);
		}
	}
#else
	for (int j = lane; j < n_stage; j += 32) {
		if (base + j < P.contrib_cap) {
			const uint2 cd   = W.stage[j];
			P.flat[base + j] = make_uint4((unsigned)W.qid[cd.x], cd.y, (unsigned)W.qenv[cd.x], 0u);
		} else {
			atomicOr(io.flags, 8);
		}
	}
#endif
	__syncwarp();
	n_stage = 0;
}

// PRISM (soft-rigid, trees of PRISM_MIN_TREE tets and more): the leaf filter also tests the three planes through the
// triangle's edges along its normal.  Measured (scripts/r02_run8.sh): C5 x 1024 (131 072-tet pads) broadphase 15.4 -> 17.8 ms,
// narrowphase 19.9 -> 15.0 ms (93.3 k -> 66.5 k candidates per env reach the clipper): step 42.5 -> 38.1 ms; C1 x 4096
// (128 tets) +3.3 / -4.1 us: a wash, left off there.
template <bool QTET, bool SWEEP, bool PRISM>
__global__ void __launch_bounds__(BP_BLOCK, FT_CTAS_PER_SM) bp_traverse_kernel(PairDesc P, StepIO io, int fixed_slots)
{
	typedef typename std::conditional<QTET, FlatQueuesSoft, FlatQueuesRigid>::type Queues;
	constexpr int NODE_CAP = Queues::node_cap, STAGE_CAP = Queues::stage_cap;
	Queues &W              = reinterpret_cast<Queues *>(bp_smem)[threadIdx.x >> 5];
	const int lane         = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	const float4 *nodes4   = reinterpret_cast<const float4 *>(P.A.nodes);
	pdl_release(); // the pair's narrowphase may become resident while this grid drains (it waits for our completion)
	pdl_wait();    // the prepare grid has completed: alive list, its count and the context blocks are visible
	int n_alive = P.counters[3];
	if (n_alive > P.alive_cap)
		n_alive = P.alive_cap; // (the prepare kernel raised the flag)
	// batch size: 16 alive queries per warp, 8 when that would leave resident warps of the grid without a batch.
	// Measured (scripts/r02_run3.sh, broadphase stage): C1 x 4096 46.4 / 44.6 / 55.4 us, C3 x 4096 0.670 / 0.576 / 0.563 ms
	// for 8 / 16 / 32: full batches fill the lanes of an iteration but leave fewer warps.  Large trees (the 131 072-tet pads
	// of config 5) are the other regime: one query alone keeps thousands of items in flight, so its frontier fills the lanes,
	// and a batch of several such queries is a long serial tail for the warp that draws it: C5 x 1024 broadphase stage
	// 10.1 / 11.3 / 14.0 / 15.5 / 18.1 ms for 1 / 2 / 4 / 8 / 16 queries per batch (scripts/r02_run17.sh).
	// Subtree split (large trees that have a split table, GeomDev::split_nodes): a work item is (alive query, one of K
	// subtrees at depth SPLIT_DEPTH), the walk starts at that subtree's root.  One query of config 5 meets thousands of
	// tets: K items of 1 / K the size balance the warps (and give a SINGLE environment K times the parallelism).
	const int total_warps = (int)gridDim.x * BP_WARPS;
	const int K           = (!SWEEP && P.A.split_nodes && P.A.env_stride == 0) ? P.A.n_split : 1;
	const long n_items    = (long)n_alive * K;
	// (small batches, down to the reference's single mjData: fewer queries per batch, one per warp when the grid has a warp
	// for every alive query: the walk of one query is ~9 dependent iterations instead of ~27 for a batch of 14;
	// traversal of config 1 with ONE environment 12.7 -> see profiles/r02_notes.md)
	// (tet queries: 32 per batch, C3 x 4096 broadphase 0.586 / 0.576 / 0.566 ms for 16 / 20 / 24, scripts/r02_run37.sh; triangle
	// queries of config 1: 0.0475 - 0.0481 ms for 8 ... 14 against 0.0440 for 16: the walk is bound by the instructions of
	// its iterations, fewer and fuller ones win even when a third of the grid's warps get no batch)
	int slots = P.n_tree >= PRISM_MIN_TREE ? 1 : (QTET ? 32 : 16);
	while (slots > 1 && n_alive < (slots / 2) * total_warps) // 16 from 8 alive queries per warp of the grid on, ... 1 below one
		slots >>= 1;
	if (K > 1) { // C5 x 1024 broadphase stage 10.8 / 9.5 / 8.7 / 9.1 ms for 4 / 8 / 16 / 32 items per batch (10.1 unsplit);
		         // small batches (one environment: 3840 items) get smaller batches so that every resident warp has one
		slots = 16;
		while (slots > 1 && n_items / slots < 2L * total_warps)
			slots >>= 1;
	}
	if (fixed_slots > 0) // tuning override (HCS_FT_SLOTS)
		slots = fixed_slots;
	const long n_batches = (n_items + slots - 1) / slots;

	// (a warp's first batch is its own index in the grid, later ones come from the work counter: narrow_kernel)
	bool first = true;
	for (;;) {
		int batch = (int)blockIdx.x * BP_WARPS + (int)(threadIdx.x >> 5);
		if (!first) {
			if (n_batches <= (long)total_warps)
				break;
			if (lane == 0)
				batch = atomicAdd(P.counters + 2, 1);
			batch = total_warps + __shfl_sync(FULL_MASK, batch, 0);
		}
		first = false;
		if ((long)batch >= n_batches)
			break;
		const long first  = (long)batch * slots;
		const int n_slots = (int)min((long)slots, n_items - first);
		int start         = 0; // where the slot's walk starts: the root, or the root of its subtree (< 0: a single leaf ~start)
		if (lane < n_slots) {
			const long item = first + lane;
			const long qi   = item / K;
			if (K > 1)
				start = P.A.split_nodes[(int)(item - qi * K)];
			const F8 *src = reinterpret_cast<const F8 *>(P.alive) + (size_t)qi * (ALIVE_WORDS / 8);
			float rec[ALIVE_WORDS];
#pragma unroll
			for (int k = 0; k < ALIVE_WORDS / 8; ++k) {
				const F8 t = ld8f_stream(src + k);
#pragma unroll
				for (int j = 0; j < 8; ++j)
					rec[8 * k + j] = t.a[j];
			}
#pragma unroll
			for (int k = 0; k < 6; ++k)
				W.qbox[k][lane] = rec[AR_BOX + k];
#pragma unroll
			for (int k = 0; k < 4; ++k)
				W.qpl[k][lane] = rec[AR_PL + k];
			constexpr int nf = QTET ? 12 : 9;
#pragma unroll
			for (int k = 0; k < nf; ++k)
				W.qvf[k][lane] = rec[AR_VF + k];
			W.qm[lane] = rec[AR_M];
			if (QTET) {
#pragma unroll
				for (int k = 0; k < 4; ++k)
					W.qx[k][lane] = rec[AR_QX + k];
			}
			W.qid[lane]  = __float_as_int(rec[AR_QID]);
			W.qenv[lane] = __float_as_int(rec[AR_ENV]);
			W.qev[lane]  = 0;
		}
		// start items: internal nodes go to the node queue, subtrees that are a single leaf to the leaf queue
		int n_stage = 0, n_leaf = 0, n_node = 0;
		if (!SWEEP) {
			const bool s_node = lane < n_slots && start >= 0, s_leaf = lane < n_slots && start < 0;
			const unsigned m_n = __ballot_sync(FULL_MASK, s_node), m_l = __ballot_sync(FULL_MASK, s_leaf);
			if (s_node)
				W.nodeq[__popc(m_n & lt_mask)] = ((unsigned)lane << ITEM_SHIFT) | (unsigned)start;
			if (s_leaf)
				W.leafq[__popc(m_l & lt_mask)] = ((unsigned)lane << ITEM_SHIFT) | (unsigned)~start;
			n_node = __popc(m_n), n_leaf = __popc(m_l);
		}
		__syncwarp();
		int sw_s = 0, sw_t = 0; // sweep position: slot, first tet of the next pass
		const int sw_slots = SWEEP ? n_slots : 0;

#pragma unroll 1
		while (n_node > 0 || n_leaf > 0 || sw_s < sw_slots) {
			if (n_leaf >= 32 || (n_node == 0 && sw_s >= sw_slots)) {
				// ---- drain up to 32 leaf items: leaf test, survivors are staged ----
				const int k = min(32, n_leaf);
				bool keep   = false;
				int skip    = 0, s = 0;
				unsigned tet = 0;
				if (lane < k) {
					const unsigned raw = W.leafq[n_leaf - k + lane];
					s = (int)(raw >> ITEM_SHIFT), tet = raw & ITEM_MASK;
					atomicAdd(&W.qev[s], 1);
					if constexpr (!QTET)
						keep = leaf_test_rigid<true, PRISM>(P, W, s, (int)tet, skip);
					else
						keep = leaf_test_soft<true>(P, W, s, (int)tet);
				}
				n_leaf -= k;
				const unsigned mk_ = __ballot_sync(FULL_MASK, keep);
				if (keep)
					W.stage[n_stage + __popc(mk_ & lt_mask)] = make_uint2((unsigned)s, tet | ((unsigned)skip << CAND_MASK_SHIFT));
				n_stage += __popc(mk_);
				__syncwarp();
				if (n_stage > STAGE_CAP - 32) // the next drain may not fit
					flush_flat(P, io, W, lane, n_stage, !QTET && slots > 1 && K == 1);
			} else if (sw_s < sw_slots) {
				// ---- sweep pass (small trees): slot sw_s against the boxes of tets sw_t .. sw_t + 31 ----
				const int t = sw_t + lane;
				bool hit    = false;
				if (t < P.n_tree) {
					float qb[6];
#pragma unroll
					for (int a = 0; a < 6; ++a)
						qb[a] = W.qbox[a][sw_s];
					float pl[4] = { 0.f, 0.f, 0.f, 0.f };
					if (!QTET) {
#pragma unroll
						for (int a = 0; a < 4; ++a)
							pl[a] = W.qpl[a][sw_s];
					}
					const F8 tb = *reinterpret_cast<const F8 *>(P.A.tet_box32 + P.A.eoff(W.qenv[sw_s]) + t);
					hit = child_overlap<!QTET>(qb, pl, tb.a[0], tb.a[1], tb.a[2], tb.a[3], tb.a[4], tb.a[5]);
				}
				const unsigned mh = __ballot_sync(FULL_MASK, hit);
				if (hit)
					W.leafq[n_leaf + __popc(mh & lt_mask)] = ((unsigned)sw_s << ITEM_SHIFT) | (unsigned)t;
				n_leaf += __popc(mh);
				sw_t += 32;
				if (sw_t >= P.n_tree)
					sw_t = 0, ++sw_s;
				__syncwarp();
			} else {
				// ---- node iteration: pop k items, push <= 2k (near the capacity fewer are popped: depth-first walk) ----
				const int k = max(1, min(min(32, n_node), NODE_CAP - n_node));
				bool pushL = false, pushR = false, leafL = false, leafR = false;
				int cl = 0, cr = 0;
				unsigned s = 0;
				if (lane < k) {
					const unsigned raw = W.nodeq[n_node - k + lane];
					s                  = raw >> ITEM_SHIFT;
					float qb[6];
#pragma unroll
					for (int a = 0; a < 6; ++a)
						qb[a] = W.qbox[a][s];
					float pl[4] = { 0.f, 0.f, 0.f, 0.f };
					if (!QTET) {
#pragma unroll
						for (int a = 0; a < 4; ++a)
							pl[a] = W.qpl[a][s];
					}
					// the 64-byte node as two 256-bit loads: every lane reads another node, an LDG costs one L1 wavefront per line it
					// touches whatever its width (records.cuh), and the node loads are most of the traversal's L1 traffic
					const F8 *nd = reinterpret_cast<const F8 *>(nodes4 + 4 * (P.A.noff(W.qenv[s]) + (size_t)(raw & ITEM_MASK)));
					const F8 n0 = nd[0], n1 = nd[1];
					const float4 a = make_float4(n0.a[0], n0.a[1], n0.a[2], n0.a[3]), b = make_float4(n0.a[4], n0.a[5], n0.a[6], n0.a[7]);
					const float4 c = make_float4(n1.a[0], n1.a[1], n1.a[2], n1.a[3]), d = make_float4(n1.a[4], n1.a[5], n1.a[6], n1.a[7]);
					cl = __float_as_int(d.x), cr = __float_as_int(d.y);
					if (child_overlap<!QTET>(qb, pl, a.x, a.y, a.z, a.w, b.x, b.y)) {
						leafL = cl < 0;
						pushL = !leafL;
					}
					if (child_overlap<!QTET>(qb, pl, b.z, b.w, c.x, c.y, c.z, c.w)) {
						leafR = cr < 0;
						pushR = !leafR;
					}
				}
				n_node -= k;
				__syncwarp();
				const unsigned mL = __ballot_sync(FULL_MASK, pushL), mR = __ballot_sync(FULL_MASK, pushR);
				const unsigned lL = __ballot_sync(FULL_MASK, leafL), lR = __ballot_sync(FULL_MASK, leafR);
				const int nL = __popc(mL), nR = __popc(mR), nlL = __popc(lL), nlR = __popc(lR);
				if (n_node + nL + nR > NODE_CAP) { // only with k forced to 1 on a full queue: a tree deeper than NODE_CAP / 32
					if (lane == 0)
						atomicOr(io.flags + 1, 1);
					n_node = 0;
					n_leaf = 0;
					break;
				}
				if (pushL)
					W.nodeq[n_node + __popc(mL & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)cl;
				if (pushR)
					W.nodeq[n_node + nL + __popc(mR & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)cr;
				if (leafL)
					W.leafq[n_leaf + __popc(lL & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)~cl;
				if (leafR)
					W.leafq[n_leaf + nlL + __popc(lR & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)~cr;
				n_node += nL + nR;
				n_leaf += nlL + nlR;
				__syncwarp();
			}
		}
		if (n_stage > 0)
			flush_flat(P, io, W, lane, n_stage, !QTET && slots > 1 && K == 1);
		// pair-evals started (LBVH leaf hits), per environment
		if (lane < n_slots && W.qev[lane] > 0)
			atomicAdd(reinterpret_cast<unsigned long long *>(P.accum + (size_t)W.qenv[lane] * ACC_WORDS + ACC_NEVALS),
			          (unsigned long long)W.qev[lane]);
		__syncwarp();
	}
}

template <bool QTET, bool SWEEP, bool PRISM = false>
static void launch_flat_bp(const PairDesc &P, const StepIO &io, cudaStream_t s)
{
	typedef typename std::conditional<QTET, FlatQueuesSoft, FlatQueuesRigid>::type Queues;
	const long total = (long)io.n_env * P.nq;
	bp_prepare_kernel<QTET><<<(unsigned)((total + PREP_BLOCK - 1) / PREP_BLOCK), PREP_BLOCK, 0, s>>>(P, io);
	const int smem = (int)sizeof(Queues) * BP_WARPS;
	auto kernel    = bp_traverse_kernel<QTET, SWEEP, PRISM>;
	ensure_dynamic_smem(kernel, smem);
	// persistent warps; never more than one warp per query element
	const int grid = (int)std::max<long>(1, std::min<long>((total + BP_WARPS - 1) / BP_WARPS, (long)io.n_sms * FT_CTAS_PER_SM));
	static const int fixed_slots = getenv("HCS_FT_SLOTS") ? std::max(1, std::min(32, atoi(getenv("HCS_FT_SLOTS")))) : 0;
	launch_chained(kernel, dim3(grid), dim3(BP_BLOCK), (size_t)smem, s, P, io, fixed_slots);
}

template <int KIND, bool SWEEP>
static void launch_bp(const PairDesc &P, const StepIO &io, cudaStream_t s, int grid)
{
	typedef typename std::conditional<KIND == 1, QueuesSoft, QueuesRigid>::type Queues;
	const int smem = (int)sizeof(Queues) * BP_WARPS;
	ensure_dynamic_smem(broadphase_kernel<KIND, SWEEP>, smem);
	broadphase_kernel<KIND, SWEEP><<<grid, BP_BLOCK, smem, s>>>(P, io);
}

void launch_broadphase(const PairDesc &P, const StepIO &io, cudaStream_t s)
{
	const long units = (long)io.n_env * P.n_slices;
	if (units == 0)
		return;
	const int grid   = (int)std::max<long>(1, std::min<long>((units + BP_WARPS - 1) / BP_WARPS, (long)io.n_sms * BP_CTAS_PER_SM));
	const bool sweep = P.n_tree <= SWEEP_MAX_TREE;
	static const bool legacy = getenv("HCS_BP_LEGACY") != nullptr; // the per-unit kernel for the tree kinds (A/B measurements)
	const bool per_env = P.A.env_stride > 0 || P.B.env_stride > 0; // per-environment geometry: the flat kernels only
	if ((!legacy || per_env) && P.alive && (P.kind == PAIR_SOFT_RIGID || P.kind == PAIR_SOFT_SOFT)) {
		if (P.kind == PAIR_SOFT_RIGID)
			sweep ? launch_flat_bp<false, true>(P, io, s) :
			        (P.n_tree >= PRISM_MIN_TREE ? launch_flat_bp<false, false, true>(P, io, s) : launch_flat_bp<false, false>(P, io, s));
		else
			sweep ? launch_flat_bp<true, true>(P, io, s) : launch_flat_bp<true, false>(P, io, s);
		return;
	}
	if (P.kind == PAIR_SOFT_RIGID)
		sweep ? launch_bp<0, true>(P, io, s, grid) : launch_bp<0, false>(P, io, s, grid);
	else if (P.kind == PAIR_SOFT_SOFT)
		sweep ? launch_bp<1, true>(P, io, s, grid) : launch_bp<1, false>(P, io, s, grid);
	else if (P.kind == PAIR_SOFT_PLANE)
		launch_bp<2, false>(P, io, s, grid);
}

} // namespace hcs
