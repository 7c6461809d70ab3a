// K3 broadphase — warp-cooperative LBVH traversal with shared-memory work queues (sm_100a).
//
// Replaces Bvh<Obb,.>::Collide inside the Drake queries the reference calls at
// mujoco_contact_surfaces_plugin.cpp:284-303 (candidate generation), and hoists the two exact early-outs
// of mesh_intersection.cc into the traversal so the clipping kernel only sees pairs that need clipping:
//   * IsFaceNormalAlongPressureGradient:  ghat . (R_SR n) > cos(5 pi / 8)
//   * trivial reject: all three triangle vertices strictly outside one tet half space (then the
//     Sutherland-Hodgman clip is empty; a 1e-12 margin keeps the decision identical to the clip's).
//
// One warp owns one (env, pair, query slice).  Lane-per-query traversal left 3 of 32 lanes busy on the
// sphere-on-box scene (most query triangles die at the root), so the warp instead shares two LIFO queues
// in shared memory:  node items (query, internal node) and leaf items (query, tet).  Every iteration all
// lanes pop one item each; children / survivors are appended with warp-ballot + popc prefix sums, which
// also makes the emitted candidate order deterministic (no atomics anywhere).
#include <algorithm>

#include "dmath.cuh"
#include "hcs_internal.h"
#include "records.cuh"

namespace hcs {

#define FULL_MASK 0xffffffffu
#ifndef HCS_BP_CTAS_PER_SM // tuning sweeps (build.py HCS_NVCC_DEFS); 4 -> 127 registers, 6 -> 80 registers + spills
#define HCS_BP_CTAS_PER_SM 4
#endif
#ifndef HCS_BP_PERSISTENT
#define HCS_BP_PERSISTENT 1
#endif
// Node iterations with <= 16 items popped with two lanes per item (one per child).  Measured and off
// (scripts/sweep_r01h.sh): C1 broadphase 0.0450 -> 0.0455 ms, C5 17.9 -> 18.4 ms, C3 0.795 -> 0.760 ms; the
// iterations are bound by their dependent chain (queue pop -> node load -> ballots -> push), not by the box tests.
#ifndef HCS_BP_PAIRED_POP
#define HCS_BP_PAIRED_POP 0
#endif
#ifndef HCS_BP_NODE256
#define HCS_BP_NODE256 0
#endif
#ifndef HCS_BP_LEAF32 // soft-rigid leaf test: conservative float filter in front of the exact fp64 early-outs
#define HCS_BP_LEAF32 1
#endif
constexpr int BP_CTAS_PER_SM = HCS_BP_CTAS_PER_SM;
constexpr int BP_WARPS = 4;
constexpr int BP_BLOCK = 32 * BP_WARPS;
constexpr int NODE_Q   = 768; // LIFO of (query slot, node); grows by <= 32 per iteration
constexpr int LEAF_Q   = 128; // (query slot, tet); drained in batches of 32
constexpr int STAGE    = 256; // staged candidates per warp; flushed in whole 32-candidate chunks when nearly full

constexpr int ITEM_SHIFT = 27; // queue item = query slot (5 bits) << 27 | node or tet index (27 bits)
constexpr unsigned ITEM_MASK = (1u << ITEM_SHIFT) - 1u;

struct __align__(16) WarpQueues {
	unsigned nodeq[NODE_Q];
	unsigned leafq[LEAF_Q];
	double qv[12][32];  // transformed query vertices (tri: 9 + rotated normal 3; tet: 12), lane-interleaved
	float qbox[6][32];
	float qpl[4][32]; // soft-rigid: the query triangle's plane (unit normal, offset) in A's frame;
	                  // soft-soft: the query tet's pressure gradient in A's frame (0..2), its field at A's origin (3)
	float qvf[12][32]; // the query's vertices in A's frame, rounded to float (leaf filters)
	float qm[32];     // soft-rigid: the filter's margin for this query (4e-6 x the largest coordinate involved);
	                  // soft-soft: the largest coordinate involved
	float qx[4][32];  // soft-soft: the query tet's unit gradient in A's frame (0..2), L1 norm of its gradient (3)
	int qid[32];
	uint2 stage[STAGE]; // surviving (query, tet) candidates waiting to be appended to the pair's flat list
};
static_assert(sizeof(WarpQueues) * BP_WARPS <= 48 * 1024, "static shared memory of a broadphase CTA");

// Append the first n_flush staged candidates (a multiple of 32 except at the end of the unit) to the pair's flat
// list as one range and link the range to the unit's chain.  WHERE the range lands depends on the order in which
// warps get here; results do not: every record carries (unit, index inside the unit), contributions are read back
// per unit along the chain, in index order.  `last`: -2 no range yet, -1 the inline first range, >= 0 pool range.
__device__ __forceinline__ void flush_stage(const PairDesc &P, const StepIO &io, WarpQueues &W, int unit, int lane,
                                            int n_flush, int &n_stage, int &i0, int &last)
{
	int base = 0;
	if (lane == 0) {
		base     = atomicAdd(P.counters, n_flush);
		int4 rec = make_int4(base, n_flush, -1, 0);
		if (last == -2) {
			P.unit_range[unit] = rec;
			last               = -1;
		} else {
			int r = atomicAdd(P.counters + 3, 1);
			if (r < P.range_cap) {
				P.ranges[r] = rec;
				if (last == -1)
					P.unit_range[unit].z = r;
				else
					P.ranges[last].z = r;
				last = r;
			} else {
				atomicOr(io.flags, 8);
			}
		}
	}
	base = __shfl_sync(FULL_MASK, base, 0);
	for (int j = lane; j < n_flush; j += 32)
		if (base + j < P.contrib_cap) {
			uint2 cd         = W.stage[j];
			P.flat[base + j] = make_uint4(cd.x, cd.y, (unsigned)unit, (unsigned)(i0 + j));
		}
	__syncwarp();
	int rem    = n_stage - n_flush; // < 32: slides to the front
	uint2 keep = make_uint2(0, 0);
	if (lane < rem)
		keep = W.stage[n_flush + lane];
	__syncwarp();
	if (lane < rem)
		W.stage[lane] = keep;
	__syncwarp();
	i0 += n_flush;
	n_stage = rem;
}

// Box overlap of the query with one child of a node; for triangle queries (PLANE) the child box must also
// straddle the triangle's plane: a subtree entirely on one side of that plane cannot meet the triangle.
// The single-interior-vertex sphere tets are slivers whose boxes all overlap a large rigid triangle's box;
// the plane test prunes those whole subtrees (profiles/r01_notes.md).
template <bool PLANE>
__device__ __forceinline__ bool box_overlap(const float *q, const float *pl, float4 a, float4 b, float4 c, bool left)
{
	// node layout: llo[3] lhi[3] rlo[3] rhi[3]
	float lo0 = left ? a.x : b.z, lo1 = left ? a.y : b.w, lo2 = left ? a.z : c.x;
	float hi0 = left ? a.w : c.y, hi1 = left ? b.x : c.z, hi2 = left ? b.y : c.w;
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

// per (env, pair) context block read by the flat narrowphase (layout: hcs_internal.h)
__device__ __forceinline__ void write_pair_ctx(const PairDesc &P, const StepIO &io, int env, const Xform &X_WA,
                                               const Xform &X_WB, const Xform &X_AB, D3 p_BAo)
{
	double *cb = P.pair_ctx + (size_t)env * PAIR_CTX_DOUBLES;
	const double *velA = io.vel + ((size_t)env * io.n_geoms + P.gA) * 6;
	const double *velB = io.vel + ((size_t)env * io.n_geoms + P.gB) * 6;
#pragma unroll
	for (int i = 0; i < 9; ++i) {
		cb[i]           = X_WA.R[i];
		cb[CTX_RAB + i] = X_AB.R[i];
	}
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		cb[CTX_WA + i] = velA[i], cb[CTX_VA + i] = velA[3 + i];
		cb[CTX_WB + i] = velB[i], cb[CTX_VB + i] = velB[3 + i];
	}
	cb[CTX_XA] = X_WA.p.x, cb[CTX_XA + 1] = X_WA.p.y, cb[CTX_XA + 2] = X_WA.p.z;
	cb[CTX_XB] = X_WB.p.x, cb[CTX_XB + 1] = X_WB.p.y, cb[CTX_XB + 2] = X_WB.p.z;
	cb[CTX_PAB] = X_AB.p.x, cb[CTX_PAB + 1] = X_AB.p.y, cb[CTX_PAB + 2] = X_AB.p.z;
	cb[CTX_PBA] = p_BAo.x, cb[CTX_PBA + 1] = p_BAo.y, cb[CTX_PBA + 2] = p_BAo.z;
	// the pad of every 32-byte group too: a sector that is only partly written is completed from DRAM when it is read
	cb[CTX_WA + 3] = cb[CTX_VA + 3] = cb[CTX_XB + 3] = cb[CTX_WB + 3] = cb[CTX_VB + 3] = cb[CTX_PBA + 3] = 0.0;
}

// Soft geom against a rigid half space (the query the reference calls at plugin.cpp:298-299): every tet of the
// unit's slice is classified against the plane, the tets it cuts become the unit's candidates.  On the mixed-objects
// scene 13 % of the tets are cut: clipping them in the flat narrowphase keeps whole warps busy, where the
// one-thread-per-tet kernel ran the whole slice-and-integrate path for the few cut tets of each 32-tet group.
__device__ __forceinline__ void plane_unit(const PairDesc &P, const StepIO &io, WarpQueues &W, int warp, int lane)
{
	int env = warp / P.n_slices, slice = warp - env * P.n_slices;
	Xform X_WA = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
	Xform X_WB = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
	Xform X_AB = invert_and_compose(X_WA, X_WB);
	D3 p_BAo   = -rotT(X_AB.R, X_AB.p);
	if (slice == 0 && lane == 0)
		write_pair_ctx(P, io, env, X_WA, X_WB, X_AB, p_BAo);
	// the half space in A's frame: normal = z column of R_AB, through p_AB (mesh_half_space_intersection.cc)
	D3 n_S    = mk(X_AB.R[2], X_AB.R[5], X_AB.R[8]);
	double pd = dot(n_S, X_AB.p);
	int n_stage = 0, i0 = 0;
	int last_range = -2;
	const int q_begin = slice * P.slice_q, q_end = min(P.nq, q_begin + P.slice_q);
	const unsigned lt_mask = (1u << lane) - 1u;
	// the whole geom above the plane: nothing can be cut
	bool above = dot(n_S, mk(P.A.bound_c[0], P.A.bound_c[1], P.A.bound_c[2])) - pd > P.A.bound_r + 1e-9;
	// 32 consecutive tets = 32 x 96 bytes of vertices at a 128-byte stride.  Lane-per-record loads touch 32 lines per
	// instruction; instead the warp reads the 96 groups of 32 bytes in index order (11 lines per instruction) and
	// transposes them through shared memory (rows padded to 13 doubles: conflict-free column reads).
	constexpr int ROW = 13;
	double *rows      = reinterpret_cast<double *>(W.nodeq); // 32 x 13 doubles = 3328 B of the (unused) queues
	static_assert(sizeof(W.nodeq) + sizeof(W.leafq) >= 32 * ROW * sizeof(double), "transpose buffer");
	for (int q0 = q_begin; q0 < q_end && !above; q0 += 32) {
		const int n_here = min(32, q_end - q0);
		const double *base = reinterpret_cast<const double *>(P.A.tet_geom + q0);
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			int j = lane + 32 * k, r = j / 3, part = j - 3 * r; // group j = (record r, 32-byte part)
			if (r < n_here) {
				D4 gq     = ld4(base + 16 * r + 4 * part);
				double *o = rows + ROW * r + 4 * part;
				o[0] = gq.x, o[1] = gq.y, o[2] = gq.z, o[3] = gq.w;
			}
		}
		__syncwarp();
		int t     = q0 + lane;
		bool keep = false;
		if (t < q_end) {
			const double *gv = rows + ROW * lane;
			int code = 0;
#pragma unroll
			for (int k = 0; k < 4; ++k)
				if (dot(n_S, mk(gv[3 * k], gv[3 * k + 1], gv[3 * k + 2])) - pd > 0)
					code |= 1 << k;
			keep = code != 0 && code != 15;
		}
		__syncwarp();
		unsigned mk_ = __ballot_sync(FULL_MASK, keep);
		if (keep)
			W.stage[n_stage + __popc(mk_ & lt_mask)] = make_uint2(0u, (unsigned)t);
		n_stage += __popc(mk_);
		__syncwarp();
		if (n_stage > STAGE - 32)
			flush_stage(P, io, W, warp, lane, n_stage & ~31, n_stage, i0, last_range);
	}
	if (n_stage > 0)
		flush_stage(P, io, W, warp, lane, n_stage, n_stage, i0, last_range);
	if (lane == 0) {
		P.unit_count[warp] = i0;
		P.unit_evals[warp] = q_end - q_begin; // tets examined
		if (last_range == -2)
			P.unit_range[warp] = make_int4(0, 0, -1, 0);
	}
}

// records the leaf test of tet t will gather: TetField (192 bytes, two lines; soft-soft reads only the second: gradient
// and ghat) and the tet's vertices
template <bool QTET>
__device__ __forceinline__ void prefetch_leaf(const PairDesc &P, int t)
{
#if HCS_BP_PREFETCH
	const char *tf = reinterpret_cast<const char *>(P.A.tet_field + t);
	if (!QTET)
		asm volatile("prefetch.global.L1 [%0];" ::"l"(tf));
	asm volatile("prefetch.global.L1 [%0];" ::"l"(tf + 128));
	asm volatile("prefetch.global.L1 [%0];" ::"l"(P.A.tet_geom + t));
#endif
}

// QTET: query elements are tets of B (soft-soft), otherwise triangles of B (soft-rigid)
template <bool QTET>
__device__ __forceinline__ void broadphase_unit(const PairDesc &P, const StepIO &io, WarpQueues &W, int warp, int lane)
{
	int env = warp / P.n_slices, slice = warp - env * P.n_slices;
	Xform X_WA = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gA);
	Xform X_WB = load_pose(io.xpos, io.xmat, io.n_geoms, env, P.gB);
	{ // pair-level reject on bounding spheres
		D3 ca = apply(X_WA, mk(P.A.bound_c[0], P.A.bound_c[1], P.A.bound_c[2]));
		D3 cb = apply(X_WB, mk(P.B.bound_c[0], P.B.bound_c[1], P.B.bound_c[2]));
		D3 d  = ca - cb;
		double rr = P.A.bound_r + P.B.bound_r + 1e-9;
		if (dot(d, d) > rr * rr) {
			if (lane == 0) {
				P.unit_count[warp] = 0;
				P.unit_evals[warp] = 0;
				P.unit_range[warp] = make_int4(0, 0, -1, 0);
			}
			return;
		}
	}
	Xform X_AB  = invert_and_compose(X_WA, X_WB);
	D3 p_BAo    = -rotT(X_AB.R, X_AB.p); // origin of A expressed in B (p_NMo of field_intersection.cc)
	if (slice == 0 && lane == 0)
		write_pair_ctx(P, io, env, X_WA, X_WB, X_AB, p_BAo);
	int n_stage = 0, i0 = 0, evals = 0; // warp-uniform: staged candidates, candidates already flushed, leaf hits
	int last_range = -2;                // lane 0 only
	int q_begin = slice * P.slice_q, q_end = min(P.nq, q_begin + P.slice_q);
	unsigned lt_mask = (1u << lane) - 1u;
	const float4 *nodes4 = reinterpret_cast<const float4 *>(P.A.nodes);
	// largest coordinate of A's geometry in its own frame (bounding sphere), for the float filter's margin
	const float leaf_scale = (float)(fmax(fmax(fabs(P.A.bound_c[0]), fabs(P.A.bound_c[1])), fabs(P.A.bound_c[2])) + P.A.bound_r);

	for (int q0 = q_begin; q0 < q_end; q0 += 32) {
		// ---- load + transform this chunk's queries, test against the root box, compact survivors ----
		int q = q0 + lane;
		bool alive = q < q_end;
		double v[12];
		float box[6];
		if (alive) {
			double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
			// 12 doubles = three 256-bit loads: the tet's vertices, or the triangle's vertices + normal
			const double *vp = QTET ? reinterpret_cast<const double *>(P.B.tet_geom + q) :
			                          reinterpret_cast<const double *>(P.B.tris + q);
			const D4 r0 = ld4(vp), r1 = ld4(vp + 4), r2 = ld4(vp + 8);
			const D3 qv[4] = { mk(r0.x, r0.y, r0.z), mk(r0.w, r1.x, r1.y), mk(r1.z, r1.w, r2.x), mk(r2.y, r2.z, r2.w) };
			const int nv   = QTET ? 4 : 3;
#pragma unroll
			for (int i = 0; i < nv; ++i) {
				D3 p = apply(X_AB, qv[i]);
				v[3 * i] = p.x, v[3 * i + 1] = p.y, v[3 * i + 2] = p.z;
				lo[0] = fmin(lo[0], p.x), lo[1] = fmin(lo[1], p.y), lo[2] = fmin(lo[2], p.z);
				hi[0] = fmax(hi[0], p.x), hi[1] = fmax(hi[1], p.y), hi[2] = fmax(hi[2], p.z);
			}
			if (!QTET) {
				D3 nS = rot(X_AB.R, qv[3]);
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
		unsigned m_alive = __ballot_sync(FULL_MASK, alive);
		int n_alive      = __popc(m_alive);
		if (n_alive == 0)
			continue;
		__syncwarp();
		if (alive) {
			int slot = __popc(m_alive & lt_mask);
#pragma unroll
			for (int k = 0; k < 12; ++k)
				W.qv[k][slot] = v[k];
#pragma unroll
			for (int k = 0; k < 6; ++k)
				W.qbox[k][slot] = box[k];
			if (!QTET) {
				W.qpl[0][slot] = (float)v[9], W.qpl[1][slot] = (float)v[10], W.qpl[2][slot] = (float)v[11];
				W.qpl[3][slot] = (float)(v[9] * v[0] + v[10] * v[1] + v[11] * v[2]);
				float big = leaf_scale;
#pragma unroll
				for (int k = 0; k < 9; ++k) {
					W.qvf[k][slot] = (float)v[k];
					big            = fmaxf(big, fabsf((float)v[k]));
				}
				W.qm[slot] = 4e-6f * big + 1e-30f;
			} else {
#if HCS_BP_LEAF32
				// what the float filter of the soft-soft leaf test needs of the query tet, computed once per query in
				// double: gradient and unit gradient rotated into A's frame, field value at A's origin
				const TetField *f1 = P.B.tet_field + q;
				const D4 ge1       = load_grad_e0(f1);
				const D3 g1M = rot(X_AB.R, xyz(ge1)), gh1M = rot(X_AB.R, load_ghat(f1));
				W.qpl[0][slot] = (float)g1M.x, W.qpl[1][slot] = (float)g1M.y, W.qpl[2][slot] = (float)g1M.z;
				W.qpl[3][slot] = (float)(dot(xyz(ge1), p_BAo) + ge1.w);
				W.qx[0][slot] = (float)gh1M.x, W.qx[1][slot] = (float)gh1M.y, W.qx[2][slot] = (float)gh1M.z;
				W.qx[3][slot] = (float)(fabs(g1M.x) + fabs(g1M.y) + fabs(g1M.z));
				float big = leaf_scale;
#pragma unroll
				for (int k = 0; k < 12; ++k) {
					W.qvf[k][slot] = (float)v[k];
					big            = fmaxf(big, fabsf((float)v[k]));
				}
				W.qm[slot] = big;
#endif
			}
			W.qid[slot]   = q;
			W.nodeq[slot] = (unsigned)slot << ITEM_SHIFT; // (slot, root)
		}
		__syncwarp();
		int n_node = n_alive, n_leaf = 0;

		// ---- shared-queue traversal ----
		while (n_node > 0 || n_leaf > 0) {
			if (n_leaf >= 32 || n_node == 0) {
				// drain up to 32 leaf items: exact early-outs, survivors go to the candidate slab
				int k     = min(32, n_leaf);
				bool keep = false;
				uint2 it  = make_uint2(0, 0);
				if (lane < k) {
					unsigned raw = W.leafq[n_leaf - k + lane];
					it           = make_uint2(raw >> ITEM_SHIFT, raw & ITEM_MASK);
					keep         = true;
					if (!QTET) {
						const TetField *tf = P.A.tet_field + it.y;
						int s  = (int)it.x;
#if HCS_BP_LEAF32
						// Conservative float filter.  The exact tests below are early-outs: a pair they reject clips to
						// nothing, so rejecting a SUBSET of those pairs here changes no result, and a pair that slips
						// through is clipped (to nothing) by the narrowphase.  One 128-byte TetLeaf32 line replaces the
						// 192 + 96 bytes of fp64 records per leaf hit; the margin qm = 4e-6 x (largest coordinate
						// involved) is > 5x the worst rounding of the float dot products and of the inputs' conversion,
						// so "outside by more than qm in float" implies "outside by more than 1e-12 in double".  NaNs
						// (degenerate tets) compare false and fall through to the exact path.  Only the gradient cull
						// is a semantic filter, not an early-out: it is decided in float only when it is clear by 1e-5.
						const TetLeaf32 *tl = P.A.tet_leaf32 + it.y;
						const F8 l0 = ld8f(tl, 0), l1 = ld8f(tl, 1), l2 = ld8f(tl, 2), l3 = ld8f(tl, 3);
						const float nx = W.qpl[0][s], ny = W.qpl[1][s], nz = W.qpl[2][s], dtf = W.qpl[3][s], m = W.qm[s];
						const float cosg = fdot3(l2.a[0], l2.a[1], l2.a[2], nx, ny, nz);
						if (cosg < (float)HCS_COS_ALPHA - 1e-5f)
							keep = false;
						else if (cosg < (float)HCS_COS_ALPHA + 1e-5f)
							keep = dot(load_ghat(tf), mk(W.qv[9][s], W.qv[10][s], W.qv[11][s])) > HCS_COS_ALPHA;
						if (keep) {
							const float ax = W.qvf[0][s], ay = W.qvf[1][s], az = W.qvf[2][s], bx = W.qvf[3][s], by = W.qvf[4][s],
							            bz = W.qvf[5][s], cx = W.qvf[6][s], cy = W.qvf[7][s], cz = W.qvf[8][s];
							const float pl[4][4] = { { l0.a[0], l0.a[1], l0.a[2], l0.a[3] }, { l0.a[4], l0.a[5], l0.a[6], l0.a[7] },
								                     { l1.a[0], l1.a[1], l1.a[2], l1.a[3] }, { l1.a[4], l1.a[5], l1.a[6], l1.a[7] } };
#pragma unroll
							for (int f = 0; f < 4; ++f) {
								float sa = fdot3(pl[f][0], pl[f][1], pl[f][2], ax, ay, az) - pl[f][3];
								float sb = fdot3(pl[f][0], pl[f][1], pl[f][2], bx, by, bz) - pl[f][3];
								float sc = fdot3(pl[f][0], pl[f][1], pl[f][2], cx, cy, cz) - pl[f][3];
								if (sa > m && sb > m && sc > m)
									keep = false;
							}
							// tet vertices: l2.a[4..7], l3.a[0..7]
							float h0 = fdot3(nx, ny, nz, l2.a[4], l2.a[5], l2.a[6]) - dtf;
							float h1 = fdot3(nx, ny, nz, l2.a[7], l3.a[0], l3.a[1]) - dtf;
							float h2 = fdot3(nx, ny, nz, l3.a[2], l3.a[3], l3.a[4]) - dtf;
							float h3 = fdot3(nx, ny, nz, l3.a[5], l3.a[6], l3.a[7]) - dtf;
							if ((h0 > m && h1 > m && h2 > m && h3 > m) || (h0 < -m && h1 < -m && h2 < -m && h3 < -m))
								keep = false;
						}
#else
						D3 nS  = mk(W.qv[9][s], W.qv[10][s], W.qv[11][s]);
						keep   = dot(load_ghat(tf), nS) > HCS_COS_ALPHA;
						if (keep) {
							D3 a = mk(W.qv[0][s], W.qv[1][s], W.qv[2][s]), b = mk(W.qv[3][s], W.qv[4][s], W.qv[5][s]),
							   c = mk(W.qv[6][s], W.qv[7][s], W.qv[8][s]);
#pragma unroll
							for (int f = 0; f < 4; ++f) {
								const D4 pl = load_plane(tf, f);
								D3 nh       = xyz(pl);
								double d    = pl.w;
								double sa = dot(nh, a) - d, sb = dot(nh, b) - d, sc = dot(nh, c) - d;
								if (sa > 1e-12 && sb > 1e-12 && sc > 1e-12)
									keep = false;
							}
							// the triangle's plane must cut the tet: all four tet vertices strictly on one side => empty
							const TetVerts tg = load_tet_verts(P.A.tet_geom + it.y);
							double dt = dot(nS, a);
							double h0 = dot(nS, tg.v0) - dt, h1 = dot(nS, tg.v1) - dt, h2 = dot(nS, tg.v2) - dt,
							       h3 = dot(nS, tg.v3) - dt;
							if ((h0 > 1e-12 && h1 > 1e-12 && h2 > 1e-12 && h3 > 1e-12) ||
							    (h0 < -1e-12 && h1 < -1e-12 && h2 < -1e-12 && h3 < -1e-12))
								keep = false;
						}
#endif
					} else {
						bool need_exact = true;
#if HCS_BP_LEAF32
						// Float filter (see the soft-rigid one above): equal-pressure plane n = grad0 - grad1, offset
						// -(e0 - f1(Mo)) / |n| in float with explicit error bounds.  en bounds the direction error of the
						// unit normal (the gradients cancel: relative error ~ eps (|grad0| + |grad1|) / |n|), epd the
						// error of the offset; both blow up when the gradients nearly cancel, and then nothing is
						// rejected here.  The two gradient culls are semantic filters: they are decided in float only
						// when clear by en + 1e-5, otherwise the pair takes the exact path below.  Every comparison is
						// written so that a NaN or an infinity leads to the exact path.
						{
							const int s = (int)it.x;
							const TetLeafSS32 *tl = P.A.tet_leafss32 + it.y;
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
#endif
						if (need_exact) {
						// soft-soft: CalcEquilibriumPlane + the two IsPlaneNormalAlongPressureGradient culls of
						// field_intersection.cc (same arithmetic as the clip kernel), then: the equal-pressure plane
						// must cut BOTH tets, otherwise slice-and-clip is empty.
						const TetField *f0 = P.A.tet_field + it.y, *f1 = P.B.tet_field + W.qid[it.x];
						int s      = (int)it.x;
						const D4 ge0 = load_grad_e0(f0), ge1 = load_grad_e0(f1);
						D3 grad0 = xyz(ge0), grad1_N = xyz(ge1);
						D3 grad1_M   = rot(X_AB.R, grad1_N);
						double f1_Mo = dot(grad1_N, p_BAo) + ge1.w;
						D3 n_M       = grad0 - grad1_M;
						double mag   = sqrt(dot(n_M, n_M));
						keep         = mag > 0.0;
						if (keep) {
							D3 nhat   = n_M / mag;
							D3 p_MQ   = -((ge0.w - f1_Mo) / mag) * nhat;
							double pd = dot(nhat, p_MQ);
							keep      = dot(nhat, load_ghat(f0)) > HCS_COS_ALPHA;
							if (keep)
								keep = dot(rotT(X_AB.R, -nhat), load_ghat(f1)) > HCS_COS_ALPHA;
							if (keep) {
								const TetVerts tg = load_tet_verts(P.A.tet_geom + it.y);
								double h0 = dot(nhat, tg.v0) - pd, h1 = dot(nhat, tg.v1) - pd, h2 = dot(nhat, tg.v2) - pd,
								       h3 = dot(nhat, tg.v3) - pd;
								if ((h0 > 1e-12 && h1 > 1e-12 && h2 > 1e-12 && h3 > 1e-12) ||
								    (h0 < -1e-12 && h1 < -1e-12 && h2 < -1e-12 && h3 < -1e-12))
									keep = false;
								double g0 = dot(nhat, mk(W.qv[0][s], W.qv[1][s], W.qv[2][s])) - pd,
								       g1 = dot(nhat, mk(W.qv[3][s], W.qv[4][s], W.qv[5][s])) - pd,
								       g2 = dot(nhat, mk(W.qv[6][s], W.qv[7][s], W.qv[8][s])) - pd,
								       g3 = dot(nhat, mk(W.qv[9][s], W.qv[10][s], W.qv[11][s])) - pd;
								if ((g0 > 1e-12 && g1 > 1e-12 && g2 > 1e-12 && g3 > 1e-12) ||
								    (g0 < -1e-12 && g1 < -1e-12 && g2 < -1e-12 && g3 < -1e-12))
									keep = false;
							}
						}
						} // need_exact
					}
				}
				n_leaf -= k;
				evals += k;
				unsigned mk_ = __ballot_sync(FULL_MASK, keep);
				if (keep)
					W.stage[n_stage + __popc(mk_ & lt_mask)] = make_uint2((unsigned)W.qid[it.x], it.y);
				n_stage += __popc(mk_);
				__syncwarp();
				if (n_stage > STAGE - 32) // the next drain may not fit: append the whole chunks
					flush_stage(P, io, W, warp, lane, n_stage & ~31, n_stage, i0, last_range);
			} else {
				// pop k items, push <= 2k: the queue grows by <= k.  Near the capacity fewer items are popped,
				// which turns the LIFO into a depth-first walk whose stack stays below the tree depth.
				// A small frontier (one environment of the sphere-on-box scene offers 12 items on average) is popped with
				// TWO lanes per item, one per child: each lane runs one box test instead of two, twice as many lanes
				// work.  Which mode runs depends on n_node only, so the candidate order stays a function of the unit.
#if HCS_BP_PAIRED_POP
				const bool paired = n_node <= 16;
#else
				const bool paired = false;
#endif
				int k = paired ? n_node : max(1, min(min(32, n_node), NODE_Q - n_node));
				bool pushL = false, pushR = false, leafL = false, leafR = false;
				int cl = 0, cr = 0;
				unsigned s = 0;
				const int item = paired ? lane >> 1 : lane;
				if (item < k) {
					unsigned raw = W.nodeq[n_node - k + item];
					uint2 it     = make_uint2(raw >> ITEM_SHIFT, raw & ITEM_MASK);
					s            = it.x;
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
					const float4 *nd = nodes4 + 4 * (size_t)it.y;
#if HCS_BP_NODE256 // the 64-byte node as two 256-bit loads instead of 3 x 128 + 64 bits (round-2 candidate, unmeasured)
					const F8 n0 = reinterpret_cast<const F8 *>(nd)[0], n1 = reinterpret_cast<const F8 *>(nd)[1];
					float4 a = make_float4(n0.a[0], n0.a[1], n0.a[2], n0.a[3]), b = make_float4(n0.a[4], n0.a[5], n0.a[6], n0.a[7]);
					float4 c = make_float4(n1.a[0], n1.a[1], n1.a[2], n1.a[3]), d = make_float4(n1.a[4], n1.a[5], n1.a[6], n1.a[7]);
#else
					float4 a = nd[0], b = nd[1], c = nd[2], d = nd[3];
#endif
					cl = __float_as_int(d.x), cr = __float_as_int(d.y);
					if (paired) { // even lane: left child, odd lane: right child; reported through the "L" slots
						const bool left = !(lane & 1);
						if (!left)
							cl = cr;
						if (box_overlap<!QTET>(qb, pl, a, b, c, left)) {
							leafL = cl < 0;
							pushL = !leafL;
						}
					} else {
						if (box_overlap<!QTET>(qb, pl, a, b, c, true)) {
							leafL = cl < 0;
							pushL = !leafL;
						}
						if (box_overlap<!QTET>(qb, pl, a, b, c, false)) {
							leafR = cr < 0;
							pushR = !leafR;
						}
					}
				}
				n_node -= k;
				__syncwarp();
				unsigned mL = __ballot_sync(FULL_MASK, pushL), mR = __ballot_sync(FULL_MASK, pushR);
				unsigned lL = __ballot_sync(FULL_MASK, leafL), lR = __ballot_sync(FULL_MASK, leafR);
				int nL = __popc(mL), nR = __popc(mR), nlL = __popc(lL), nlR = __popc(lR);
				if (n_node + nL + nR > NODE_Q) { // cannot happen for trees shallower than NODE_Q/32 levels
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
				// leaf items wait in the queue for an iteration or more: their records travel meanwhile
				if (leafL) {
					W.leafq[n_leaf + __popc(lL & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)~cl;
					prefetch_leaf<QTET>(P, ~cl);
				}
				if (leafR) {
					W.leafq[n_leaf + nlL + __popc(lR & lt_mask)] = (s << ITEM_SHIFT) | (unsigned)~cr;
					prefetch_leaf<QTET>(P, ~cr);
				}
				n_node += nL + nR;
				n_leaf += nlL + nlR;
				__syncwarp();
			}
		}
	}
	if (n_stage > 0)
		flush_stage(P, io, W, warp, lane, n_stage, n_stage, i0, last_range);
	if (lane == 0) {
		P.unit_count[warp] = i0;
		P.unit_evals[warp] = evals;
		if (last_range == -2)
			P.unit_range[warp] = make_int4(0, 0, -1, 0);
	}
}

__device__ __forceinline__ int claim_unit(int32_t *counter, int lane)
{
	int u = 0;
	if (lane == 0)
		u = atomicAdd(counter, 1);
	return __shfl_sync(FULL_MASK, u, 0);
}
// Prefetches (next unit's poses, leaf records at push time) were measured and are off: the L1 data pipe is the busiest
// unit of this kernel and the extra wavefronts cost more than the latency they hide (C1 broadphase 0.050 -> 0.075 ms).
#ifndef HCS_BP_PREFETCH
#define HCS_BP_PREFETCH 0
#endif
#ifndef HCS_EARLY_CLAIM
#define HCS_EARLY_CLAIM 0
#endif
__device__ __forceinline__ void prefetch_l1(const void *p)
{
#if HCS_BP_PREFETCH
	asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
// Persistent warps: the resident CTAs of every SM pull (env, slice) units from a work counter, so the grid is
// never a fractional number of waves and a slow unit does not hold three finished warps' resources.
// KIND: 0 triangles of B against the tet tree of A, 1 tets of B against it, 2 the tets of A against a half space
template <int KIND>
__global__ void __launch_bounds__(BP_BLOCK, BP_CTAS_PER_SM) broadphase_kernel(PairDesc P, StepIO io)
{
	__shared__ WarpQueues sm[BP_WARPS];
	WarpQueues &W     = sm[threadIdx.x >> 5];
	const int lane    = threadIdx.x & 31;
	const int n_units = io.n_env * P.n_slices;
	pdl_release(); // the pair's narrowphase may become resident while this grid drains (it waits for our completion)
#if HCS_BP_PERSISTENT
	// HCS_EARLY_CLAIM=1 issues the work-counter atomic for the next unit before the current unit starts (its round
	// trip is 4 % of the stall samples); measured slower (0.0503 -> 0.0550 ms on C1) and off by default.
#if HCS_EARLY_CLAIM
	int unit = claim_unit(P.counters + 2, lane);
	while (unit < n_units) {
		const int requested = lane == 0 ? atomicAdd(P.counters + 2, 1) : 0;
		if (KIND == 2)
			plane_unit(P, io, W, unit, lane);
		else
			broadphase_unit<KIND == 1>(P, io, W, unit, lane);
		__syncwarp();
		unit = __shfl_sync(FULL_MASK, requested, 0);
	}
#else
	for (;;) {
		int unit = claim_unit(P.counters + 2, lane);
		if (unit >= n_units)
			break;
		if (KIND == 2)
			plane_unit(P, io, W, unit, lane);
		else
			broadphase_unit<KIND == 1>(P, io, W, unit, lane);
		__syncwarp();
	}
#endif
#else
	int unit = (blockIdx.x * BP_BLOCK + threadIdx.x) >> 5;
	if (unit < n_units) {
		if (KIND == 2)
			plane_unit(P, io, W, unit, lane);
		else
			broadphase_unit<KIND == 1>(P, io, W, unit, lane);
	}
#endif
}

void launch_broadphase(const PairDesc &P, const StepIO &io, cudaStream_t s)
{
	long units = (long)io.n_env * P.n_slices;
	if (units == 0)
		return;
	int grid = (int)((units + BP_WARPS - 1) / BP_WARPS);
#if HCS_BP_PERSISTENT
	grid = (int)std::min<long>(grid, (long)io.n_sms * BP_CTAS_PER_SM);
#endif
	if (P.kind == PAIR_SOFT_RIGID)
		broadphase_kernel<0><<<grid, BP_BLOCK, 0, s>>>(P, io);
	else if (P.kind == PAIR_SOFT_SOFT)
		broadphase_kernel<1><<<grid, BP_BLOCK, 0, s>>>(P, io);
	else if (P.kind == PAIR_SOFT_PLANE)
		broadphase_kernel<2><<<grid, BP_BLOCK, 0, s>>>(P, io);
}

} // namespace hcs
