// Narrowphase + finalize kernels of the hydroelastic contact engine (sm_100a, fp64, -fmad=false).
//
//   K4 narrow_kernel<0>        one thread per (tet, triangle) candidate that survived the broadphase
//                              early-outs: Sutherland-Hodgman clip against the tet's four precomputed half
//                              spaces, duplicate removal, polygon quadrature (mujoco_contact_surfaces_plugin.
//                              cpp:320-409) and the force law (plugin.cpp:411-483).  Restates Drake
//                              mesh_intersection.cc.
//   K5 narrow_kernel<2>        one thread per tet the half space cuts (classified by the broadphase): marching-tets
//                              slice (App. A.5).
//   K6 narrow_kernel<1>        one thread per (tet, tet) candidate: equal-pressure plane, slice + clip (A.6).
//   K7 finalize_kernel         exact accumulators -> per-pair results and per-geom wrenches (replaces two mj_applyFT
//                              per face, plugin.cpp:477-482).
// The candidate bodies live in narrow.cuh.  All three narrowphase kernels are flat over the batch: persistent warps pull
// 32-candidate chunks of the pair's candidate list, each lane clips ONE candidate whatever environment it belongs to.
//
// Reduction (round 2).  Round 1 wrote an 80-byte contribution per candidate and summed them one kernel later along
// per-unit chains in a fixed order (15 MB read back per step on C1, 14 % of the step).  Now every lane turns its
// contribution into two-limb fixed-point integers (hcs_internal.h "exact accumulators"), the warp adds the lanes of each
// run of equal environments through shared memory, and one lane per word adds the run's total to the (env, pair)
// accumulator with a 64-bit integer atomic.  Integer sums are exact, so the result does not depend on which warp clipped
// what in which order: bit-reproducible without ordered records, and the finalize kernel reads 208 bytes per (env, pair).
#include <algorithm>

#include "narrow.cuh"

namespace hcs {

constexpr int NP_WARPS = 4;
constexpr int NP_BLOCK = 32 * NP_WARPS;

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

__device__ __forceinline__ int next_chunk(int32_t *counter, int lane)
{
	int chunk = 0;
	if (lane == 0)
		chunk = atomicAdd(counter, 1);
	return __shfl_sync(FULL_MASK, chunk, 0);
}

// candidates in the flat list, clamped to its capacity (overflow is reported, not UB)
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

// ---- exact accumulation ----------------------------------------------------------------------------------
// v = hi * 2^-36 + lo * 2^-80 (+ less than 2^-81): hi = round(v * 2^36), lo = round((v - hi * 2^-36) * 2^80).  The
// residual is exact in double while |hi| < 2^53, i.e. |v| < 2^17; larger values keep 53 bits in hi.  The conversion is
// a function of v alone, so sums of limbs are exact integer sums of deterministic values.
__device__ __forceinline__ void to_limbs(double v, long long &hi, long long &lo)
{
	// (splitting the limbs off with two additions of 1.5 * 2^k constants instead of the 64-bit conversions was measured on
	// one box, three alternations: narrowphase 0.0438 vs 0.0440 ms on C1 x 4096: no difference, not kept)
	hi             = __double2ll_rn(v * ACC_HI_SCALE);
	const double r = fma(-__ll2double_rn(hi), 1.0 / ACC_HI_SCALE, v);
	lo             = __double2ll_rn(r * ACC_LO_SCALE);
}

// Copies of the column loop's body (with its two atomics) in accumulate_chunk: 16 (fully unrolled) for the small kernels,
// 4 for the large ones (tet-tet, and everything that also emits tactile triangles), which are 50 - 60 KB of SASS against a
// 32 KB instruction cache.  Measured (scripts/r02_run38.sh, narrowphase stage): C3 0.915 -> 0.895 ms, C5 x 512 7.64 -> 7.53 ms
// with 4; C1 0.0405 -> 0.0411, C4 0.0904 -> 0.0931 ms with 4: those keep 16.
#ifndef HCS_ACC_UNROLL_LARGE
#define HCS_ACC_UNROLL_LARGE 4
#endif
#ifndef HCS_ACC_UNROLL_SMALL
#define HCS_ACC_UNROLL_SMALL 16
#endif
constexpr int RED_PAIRS  = 11; // 10 components (hi, lo) + (packed counts, n_clipped)
constexpr int RED_STRIDE = 23; // words per candidate (22 used): odd, so that candidate rows start in distinct bank pairs
struct WarpReduce {            // [candidate][word], aliased with the (dead) polygon tile of the warp
	long long w[32 * RED_STRIDE];
};

// All 32 lanes call this once per chunk, after the last use of the polygon tile.  env < 0: the lane had no candidate.
// Lane j writes its candidate's 11 word pairs; then the two half warps each take 16 candidates: lane c < 11 of a half owns
// pair c and adds every run of equal environments of its half to that environment's accumulator (two 64-bit integer
// atomics per run).  The broadphase appends a unit's candidates together, so a chunk holds a few runs; runs that straddle
// the halves (or chunks) simply arrive as several exact additions.
template <int ACC_UNROLL>
__device__ __forceinline__ void accumulate_chunk(const PairDesc &P, const StepIO &io, WarpReduce &R, int lane, int env, const Acc &acc)
{
	__syncwarp(); // every lane is done with its polygon buffers
	const double d[10] = { acc.F.x, acc.F.y, acc.F.z, acc.tau.x, acc.tau.y, acc.tau.z, acc.area, acc.ac.x, acc.ac.y, acc.ac.z };
	long long *mine    = R.w + lane * RED_STRIDE;
	double big = 0;
#pragma unroll
	for (int k = 0; k < 10; ++k) {
		const double v = d[k] * P.acc_scale;
		big            = fmax(big, fabs(v)); // fmax drops NaN: checked below through the sum
		long long hi, lo;
		to_limbs(v, hi, lo);
		mine[2 * k]     = hi;
		mine[2 * k + 1] = lo;
	}
	const double sum = ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7])) + (d[8] + d[9]);
	if (env >= 0 && !(big < 0x1p26 && sum == sum)) // not finite, or out of the accumulators' range: reported
		atomicOr(io.flags, 16);
	mine[20] = env >= 0 ? ((long long)acc.n_faces | ((long long)acc.n_polygons << ACC_POLY_SHIFT) |
	                       (long long)((unsigned long long)acc.n_points << ACC_POINT_SHIFT)) : 0;
	mine[21] = env >= 0 ? 1 : 0;
	__syncwarp();
	const int half = lane >> 4, c = lane & 15;
	const bool own = c < RED_PAIRS;
	const long long *col = R.w + (half * 16) * RED_STRIDE + 2 * (own ? c : 0);
	long long s0 = 0, s1 = 0;
	int cur = __shfl_sync(FULL_MASK, env, half * 16);
#pragma unroll ACC_UNROLL
	for (int j = 0; j < 16; ++j) {
		const int e = __shfl_sync(FULL_MASK, env, half * 16 + j);
		if (e != cur) { // (uniform within a half warp)
			if (own && cur >= 0 && (s0 | s1) != 0) {
				unsigned long long *a = reinterpret_cast<unsigned long long *>(P.accum + (size_t)cur * ACC_WORDS + 2 * c);
				atomicAdd(a, (unsigned long long)s0);
				atomicAdd(a + 1, (unsigned long long)s1);
			}
			s0 = s1 = 0;
			cur     = e;
		}
		s0 += col[j * RED_STRIDE];
		s1 += col[j * RED_STRIDE + 1];
	}
	if (own && cur >= 0 && (s0 | s1) != 0) {
		unsigned long long *a = reinterpret_cast<unsigned long long *>(P.accum + (size_t)cur * ACC_WORDS + 2 * c);
		atomicAdd(a, (unsigned long long)s0);
		atomicAdd(a + 1, (unsigned long long)s1);
	}
	__syncwarp(); // the next chunk's polygons may overwrite the words
}

template <class TILE>
struct NpSmem { // shared memory of one narrowphase warp: the polygon tile, reused for the reduction words
	static constexpr size_t value = sizeof(TILE) > sizeof(WarpReduce) ? sizeof(TILE) : sizeof(WarpReduce);
};

// =================================================================================================
// the three flat narrowphase kernels: KIND 0 (tet, triangle), 1 (tet, tet), 2 (tet, half space)
// =================================================================================================
template <int KIND, bool TRI, class TILE, int WARPS, int CTAS>
__global__ void __launch_bounds__(32 * WARPS, CTAS) narrow_kernel(PairDesc P, StepIO io)
{
	pdl_release(); // the finalize kernel may become resident while this grid drains
	pdl_wait();    // the broadphase grid has completed: flat list, counters and context blocks are visible
	const int lane = threadIdx.x & 31;
	char *wbase    = reinterpret_cast<char *>(smem_d) + (size_t)(threadIdx.x >> 5) * NpSmem<TILE>::value;
	TILE &T        = *reinterpret_cast<TILE *>(wbase);
	WarpReduce &R  = *reinterpret_cast<WarpReduce *>(wbase);
	const unsigned buf0 = smem_addr(&T.xyz[0][0][lane]), buf_stride = (unsigned)sizeof(T.xyz) / (8u / SM_UNIT);
	const int total    = flat_total(P, io);
	const int n_chunks = (total + 31) >> 5;
	// Every warp's FIRST chunk is its own index in the grid, later ones come from the work counter: at kernel start all
	// warps of the grid (2368 on config 1) would otherwise queue up on ONE address, and a step of one environment pays
	// two L2 round trips per warp for a counter that only ever hands out chunk 0 and "done".
	const int total_warps = (int)gridDim.x * WARPS;
	int chunk             = (int)blockIdx.x * WARPS + (int)(threadIdx.x >> 5);
#pragma unroll 1
	while (chunk < n_chunks) {
		const int g = chunk * 32 + lane;
		int tfaces = 0, cur = 0, nv = 0;
		D3 cen     = mk(0, 0, 0);
		double ec  = 0;
		uint4 rec  = make_uint4(0, 0, 0, 0); // (query element of B, tree element of A | skip mask, env, -)
		const bool have = g < total;
		if (have)
#if HCS_STREAM_LOADS
			rec = __ldcs(P.flat + g); // (read once: streaming)
#else
			rec = P.flat[g];
#endif
		const int env = (int)rec.z;
		CandCtx ctx   = cand_ctx(P, io, env);
		Acc acc       = zero_acc();
		const int elemA = (int)(rec.y & CAND_ELEM_MASK), elemB = (int)rec.x;
		if (have) {
			if (KIND == 0)
				tfaces = cand_tet_tri<TRI>(P, io, ctx, elemB, elemA, (int)(rec.y >> CAND_MASK_SHIFT), buf0, buf_stride, acc, cur, cen,
				                           ec, nv);
			else if (KIND == 1)
				tfaces = cand_tet_tet<TRI>(P, io, ctx, elemB, elemA, buf0, buf_stride, acc, cur, cen, ec, nv);
			else
				tfaces = cand_tet_plane<TRI>(P, io, ctx, elemA, buf0, buf_stride, acc, cen, ec, nv);
			P.nverts[g] = (uint8_t)nv; // diagnostics: hcs_get_emitted
		}
		if (TRI && P.emit_tactile) {
			if (KIND == 2)
				emit_tactile<true>(tfaces, Poly{ buf0 }, PressTile{ buf0 + buf_stride }, cen, ec, ctx, io, lane, elemA, 0);
			else
				emit_tactile<false>(tfaces, Poly{ buf0 + cur * buf_stride }, PressTile{ buf0 + (cur ^ 1) * buf_stride }, cen, ec, ctx,
				                    io, lane, elemA, elemB);
		}
#ifndef HCS_NP_NO_ACCUM // timing experiments only: results are wrong without it
		accumulate_chunk<(KIND == 1 || TRI) ? HCS_ACC_UNROLL_LARGE : HCS_ACC_UNROLL_SMALL>(P, io, R, lane, have ? env : -1, acc);
#endif
		if (n_chunks <= total_warps)
			break; // (every chunk was some warp's first)
		chunk = total_warps + next_chunk(P.counters + 1, lane);
	}
}

// =================================================================================================
// K7 finalize: one thread per environment.  Reads (and clears) the exact accumulators of its pairs, writes the
// per-pair results, adds them per geom in pair order.
// =================================================================================================
__device__ __forceinline__ double from_limbs(long long hi, long long lo)
{
	const long long carry = lo >> 44; // lo holds up to 2^19 contributions of < 2^44 each: move whole hi units over
	hi += carry;
	lo -= carry << 44;
	return __ll2double_rn(hi) * (1.0 / ACC_HI_SCALE) + __ll2double_rn(lo) * (1.0 / ACC_LO_SCALE);
}

// One thread per environment walks its pairs: the better variant for scenes with one or two pairs (config 1: 10.4 us of
// stage time against 12.0 for the (env, pair)-parallel kernel below, which wins from three pairs on: config 4 24.6 -> 16.5 us,
// config 5 27 -> 13 us).
#ifndef HCS_FIN1_BLOCK
#define HCS_FIN1_BLOCK 32 // 128 CTAs for 4096 environments instead of 32: C1 x 4096 value 45.35 -> 45.77 M env-steps/s (scripts/r02_run36.sh)
#endif
constexpr int FIN1_BLOCK = HCS_FIN1_BLOCK;
__global__ void __launch_bounds__(FIN1_BLOCK) finalize_env_kernel(const PairDesc *pairs, StepIO io, int use_smem)
{
	extern __shared__ double fin_smem[]; // [n_geoms * 6][FIN1_BLOCK] wrench accumulators (scenes with many geoms: below)
	pdl_wait(); // chained behind the last narrowphase kernel (no-op otherwise)
	const int env = blockIdx.x * FIN1_BLOCK + threadIdx.x;
	if (env < io.n_env) {
		double *w = fin_smem + threadIdx.x;
		if (use_smem)
			for (int k = 0; k < io.n_geoms * 6; ++k)
				w[k * FIN1_BLOCK] = 0.0;
		for (int p = 0; p < io.n_pairs; ++p) {
			const PairDesc &P = pairs[p];
			hcs_pair_result r;
			for (int k = 0; k < 3; ++k)
				r.F[k] = r.tau[k] = r.centroid[k] = 0;
			r.area = 0;
			r.gM = P.gM, r.gN = P.gN;
			r.n_polygons = r.n_faces = r.n_points = r.n_candidates = r.n_clipped = r.reserved = 0;
			if (P.kind != PAIR_NONE) {
				long long *a = reinterpret_cast<long long *>(P.accum) + (size_t)env * ACC_WORDS;
				long long wds[ACC_WORDS];
				const longlong2 *a2 = reinterpret_cast<const longlong2 *>(a); // 208-byte records: 16-byte aligned
#pragma unroll
				for (int k = 0; k < ACC_WORDS / 2; ++k) {
					const longlong2 q = a2[k];
					wds[2 * k] = q.x, wds[2 * k + 1] = q.y;
				}
#pragma unroll
				for (int k = 0; k < ACC_WORDS / 2; ++k)
					reinterpret_cast<longlong2 *>(a)[k] = make_longlong2(0, 0); // zero between steps
				double d[10];
#pragma unroll
				for (int k = 0; k < 10; ++k)
					d[k] = from_limbs(wds[2 * k], wds[2 * k + 1]) * P.acc_unscale;
				for (int k = 0; k < 3; ++k) {
					r.F[k]        = P.sign * d[k];
					r.tau[k]      = P.sign * d[3 + k];
					r.centroid[k] = d[6] > 0 ? d[7 + k] / d[6] : 0.0;
				}
				r.area = d[6];
				const unsigned long long cnt = (unsigned long long)wds[ACC_COUNTS];
				r.n_faces      = (int)(cnt & ((1ull << ACC_POLY_SHIFT) - 1));
				r.n_polygons   = (int)((cnt >> ACC_POLY_SHIFT) & ((1ull << (ACC_POINT_SHIFT - ACC_POLY_SHIFT)) - 1));
				r.n_points     = (int)(cnt >> ACC_POINT_SHIFT);
				r.n_clipped    = (int)wds[ACC_NCLIPPED];
				r.n_candidates = (int)wds[ACC_NEVALS];
				if (use_smem)
					for (int k = 0; k < 3; ++k) {
						w[(6 * r.gM + k) * FIN1_BLOCK] += r.F[k], w[(6 * r.gM + 3 + k) * FIN1_BLOCK] += r.tau[k];
						w[(6 * r.gN + k) * FIN1_BLOCK] -= r.F[k], w[(6 * r.gN + 3 + k) * FIN1_BLOCK] -= r.tau[k];
					}
			}
			io.pair_out[(size_t)env * io.n_pairs + p] = r;
		}
		double *out = io.geom_wrench + (size_t)env * io.n_geoms * 6;
		if (!use_smem) { // too many geoms for a shared accumulator per thread: per geom, its pairs in pair order (same sums)
			for (int g = 0; g < io.n_geoms; ++g) {
				double wg[6] = { 0, 0, 0, 0, 0, 0 };
				for (int p = 0; p < io.n_pairs; ++p) {
					if (pairs[p].kind == PAIR_NONE)
						continue;
					const hcs_pair_result &r = io.pair_out[(size_t)env * io.n_pairs + p];
					if (r.gM == g)
						for (int k = 0; k < 3; ++k)
							wg[k] += r.F[k], wg[3 + k] += r.tau[k];
					if (r.gN == g)
						for (int k = 0; k < 3; ++k)
							wg[k] -= r.F[k], wg[3 + k] -= r.tau[k];
				}
				for (int k = 0; k < 6; ++k)
					out[6 * g + k] = wg[k];
				if (io.geom_wrench_host)
					for (int k = 0; k < 6; ++k)
						io.geom_wrench_host[((size_t)env * io.n_geoms + g) * 6 + k] = wg[k];
			}
		}
	}
	if (use_smem) {
		// the block's wrenches are contiguous in the output: written by consecutive threads (the end-to-end path writes the
		// caller's copy straight into mapped pinned memory: coalesced posted writes instead of 8-byte ones)
		__syncthreads();
		const int env0 = blockIdx.x * FIN1_BLOCK, n_here = min(FIN1_BLOCK, io.n_env - env0), per = io.n_geoms * 6;
		for (int i = threadIdx.x; i < n_here * per; i += FIN1_BLOCK) {
			const int e = i / per, k = i - e * per;
			const double v = fin_smem[k * FIN1_BLOCK + e];
			io.geom_wrench[(size_t)env0 * per + i] = v;
			if (io.geom_wrench_host)
				io.geom_wrench_host[(size_t)env0 * per + i] = v;
		}
	}
	// The finalize kernel is the last kernel of a step without sensors: one thread mirrors the error flags into the
	// caller's mapped pinned memory, so that the end-to-end path needs no copy after the kernels (every kernel that can
	// raise a flag has finished: stream order).
	if (io.flags_host && blockIdx.x == 0 && threadIdx.x == 0)
		for (int k = 0; k < 4; ++k)
			io.flags_host[k] = io.flags[k];
	if (io.zero_next) // the other set of step counters, for the next step (hcs_internal.h StepIO::zero_next)
		for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < io.n_zero; i += gridDim.x * blockDim.x)
			io.zero_next[i] = 0;
}

// One work item per (environment, pair): read (and clear) the pair's exact accumulators, write hcs_pair_result; then one
// item per (environment, geom): the geom's wrench = its pairs' results in pair order.  A block covers whole environments
// (FIN_BLOCK / n_pairs of them, at least one), so the second phase finds the first phase's results in shared memory.
// Round 1/early round 2 ran ONE thread per environment through all its pairs (ncu: 9.5 us for the 4096 environments of
// config 1 on 32 CTAs, 24.8 us for config 4 with four pairs: a chain of dependent loads per pair).
constexpr int FIN_BLOCK = 64;
__global__ void __launch_bounds__(FIN_BLOCK) finalize_kernel(const PairDesc *pairs, StepIO io, int envs_per_block)
{
	extern __shared__ double fin_smem[]; // [envs_per_block][n_pairs][6]: F, tau of every (env, pair) of the block
	pdl_wait(); // chained behind the last narrowphase kernel (no-op otherwise)
	const int env0 = blockIdx.x * envs_per_block, n_here = min(envs_per_block, io.n_env - env0);
	const int np = io.n_pairs, ng = io.n_geoms;
	for (int it = threadIdx.x; it < n_here * np; it += FIN_BLOCK) {
		const int el = it / np, p = it - el * np, env = env0 + el;
		const PairDesc &P = pairs[p];
		hcs_pair_result r;
		for (int k = 0; k < 3; ++k)
			r.F[k] = r.tau[k] = r.centroid[k] = 0;
		r.area = 0;
		r.gM = P.gM, r.gN = P.gN;
		r.n_polygons = r.n_faces = r.n_points = r.n_candidates = r.n_clipped = r.reserved = 0;
		if (P.kind != PAIR_NONE) {
			long long *a = reinterpret_cast<long long *>(P.accum) + (size_t)env * ACC_WORDS;
			long long wds[ACC_WORDS];
			const longlong2 *a2 = reinterpret_cast<const longlong2 *>(a); // 192-byte records: 16-byte aligned
#pragma unroll
			for (int k = 0; k < ACC_WORDS / 2; ++k) {
				const longlong2 q = a2[k];
				wds[2 * k] = q.x, wds[2 * k + 1] = q.y;
			}
#pragma unroll
			for (int k = 0; k < ACC_WORDS / 2; ++k)
				reinterpret_cast<longlong2 *>(a)[k] = make_longlong2(0, 0); // zero between steps
			double d[10];
#pragma unroll
			for (int k = 0; k < 10; ++k)
				d[k] = from_limbs(wds[2 * k], wds[2 * k + 1]) * P.acc_unscale;
			for (int k = 0; k < 3; ++k) {
				r.F[k]        = P.sign * d[k];
				r.tau[k]      = P.sign * d[3 + k];
				r.centroid[k] = d[6] > 0 ? d[7 + k] / d[6] : 0.0;
			}
			r.area = d[6];
			const unsigned long long cnt = (unsigned long long)wds[ACC_COUNTS];
			r.n_faces      = (int)(cnt & ((1ull << ACC_POLY_SHIFT) - 1));
			r.n_polygons   = (int)((cnt >> ACC_POLY_SHIFT) & ((1ull << (ACC_POINT_SHIFT - ACC_POLY_SHIFT)) - 1));
			r.n_points     = (int)(cnt >> ACC_POINT_SHIFT);
			r.n_clipped    = (int)wds[ACC_NCLIPPED];
			r.n_candidates = (int)wds[ACC_NEVALS];
		}
		double *w = fin_smem + (size_t)it * 6;
		for (int k = 0; k < 3; ++k)
			w[k] = r.F[k], w[3 + k] = r.tau[k];
		io.pair_out[(size_t)env * np + p] = r;
	}
	__syncthreads();
	// per-geom wrenches: the geom's pairs in pair order (+ as M, - as N); the block's wrenches are contiguous in the output
	// (the end-to-end path writes the caller's copy straight into mapped pinned memory)
	for (int it = threadIdx.x; it < n_here * ng; it += FIN_BLOCK) {
		const int el = it / ng, g = it - el * ng;
		double wg[6] = { 0, 0, 0, 0, 0, 0 };
		for (int p = 0; p < np; ++p) {
			const PairDesc &P = pairs[p];
			if (P.kind == PAIR_NONE)
				continue;
			const double *w = fin_smem + ((size_t)el * np + p) * 6;
			if (P.gM == g)
				for (int k = 0; k < 6; ++k)
					wg[k] += w[k];
			if (P.gN == g)
				for (int k = 0; k < 6; ++k)
					wg[k] -= w[k];
		}
		const size_t o = ((size_t)(env0 + el) * ng + g) * 6;
		for (int k = 0; k < 6; ++k)
			io.geom_wrench[o + k] = wg[k];
		if (io.geom_wrench_host)
			for (int k = 0; k < 6; ++k)
				io.geom_wrench_host[o + k] = wg[k];
	}
	// The finalize kernel is the last kernel of a step without sensors: one thread mirrors the error flags into the
	// caller's mapped pinned memory, so that the end-to-end path needs no copy after the kernels (every kernel that can
	// raise a flag has finished: stream order).
	if (io.flags_host && blockIdx.x == 0 && threadIdx.x == 0)
		for (int k = 0; k < 4; ++k)
			io.flags_host[k] = io.flags[k];
	if (io.zero_next) // the other set of step counters, for the next step (hcs_internal.h StepIO::zero_next)
		for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < io.n_zero; i += gridDim.x * blockDim.x)
			io.zero_next[i] = 0;
}

// =================================================================================================
// launchers
// =================================================================================================
template <int KIND, bool TRI, class TILE, int WARPS, int CTAS>
static void launch_np(int grid, const PairDesc &P, const StepIO &io, cudaStream_t s, bool chained)
{
	// opt in to > 48 KB dynamic shared memory (idempotent and cheap; contexts may live on several devices)
	const int smem = (int)NpSmem<TILE>::value * WARPS;
	auto kernel    = narrow_kernel<KIND, TRI, TILE, WARPS, CTAS>;
	ensure_dynamic_smem(kernel, smem);
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
				launch_np<0, true, WarpTile<7, 6>, HCS_NP_TRI_WARPS, HCS_NP_TRI_CTAS>(flat_grid(HCS_NP_TRI_CTAS, HCS_NP_TRI_WARPS), P, io, s, chained);
			else
				launch_np<0, false, WarpTile<7, 6>, HCS_NP_TRI_WARPS, HCS_NP_TRI_CTAS>(flat_grid(HCS_NP_TRI_CTAS, HCS_NP_TRI_WARPS), P, io, s, chained);
			break;
		case PAIR_SOFT_SOFT:
			if (tri)
				launch_np<1, true, WarpTile<8, 7>, NP_WARPS, HCS_NP_TET_CTAS>(flat_grid(HCS_NP_TET_CTAS), P, io, s, chained);
			else
				launch_np<1, false, WarpTile<8, 7>, NP_WARPS, HCS_NP_TET_CTAS>(flat_grid(HCS_NP_TET_CTAS), P, io, s, chained);
			break;
		case PAIR_SOFT_PLANE:
			if (tri)
				launch_np<2, true, WarpTile<4, 2>, NP_WARPS, HCS_NP_PLANE_CTAS>(flat_grid(HCS_NP_PLANE_CTAS), P, io, s, chained);
			else
				launch_np<2, false, WarpTile<4, 2>, NP_WARPS, HCS_NP_PLANE_CTAS>(flat_grid(HCS_NP_PLANE_CTAS), P, io, s, chained);
			break;
		default:
			break;
	}
}

int launch_finalize(const PairDesc *d_pairs, const StepIO &io, cudaStream_t s, bool chained)
{
	if (io.n_env <= 0)
		return 0;
	const int np = std::max(io.n_pairs, 1);
	if (np <= 2) { // one thread per environment
		const int grid1 = (io.n_env + FIN1_BLOCK - 1) / FIN1_BLOCK;
		size_t smem1    = (size_t)FIN1_BLOCK * io.n_geoms * 6 * sizeof(double);
		const int use_smem = smem1 <= 96 * 1024;
		if (!use_smem)
			smem1 = 0;
		ensure_dynamic_smem(finalize_env_kernel, (int)std::max<size_t>(smem1, 1024));
		if (chained)
			launch_chained(finalize_env_kernel, dim3(grid1), dim3(FIN1_BLOCK), smem1, s, d_pairs, io, use_smem);
		else
			finalize_env_kernel<<<grid1, FIN1_BLOCK, smem1, s>>>(d_pairs, io, use_smem);
		return 1;
	}
	// whole environments per block; the block's (env, pair) wrenches live in shared memory (<= 48 KB: up to 1024 pairs)
	const int envs_per_block = std::max(1, FIN_BLOCK / np);
	const int grid           = (io.n_env + envs_per_block - 1) / envs_per_block;
	const size_t smem        = (size_t)envs_per_block * np * 6 * sizeof(double);
	ensure_dynamic_smem(finalize_kernel, (int)std::max<size_t>(smem, 1024));
	if (chained)
		launch_chained(finalize_kernel, dim3(grid), dim3(FIN_BLOCK), smem, s, d_pairs, io, envs_per_block);
	else
		finalize_kernel<<<grid, FIN_BLOCK, smem, s>>>(d_pairs, io, envs_per_block);
	return 1;
}

} // namespace hcs
