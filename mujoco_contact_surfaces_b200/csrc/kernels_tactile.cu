// K8 tactile rasterisation (Myrmex flat taxel array), float32 like the reference.
//
// Replaces FlatTactileSensor::bvh_update (mujoco_contact_surface_sensors/src/flat_tactile_sensor.cpp:
// 262-402) and the per-update BLAS/TLAS rebuild + traversal of src/bvh.cpp:90-476.  The reference
// casts sampling_resolution^2 parallel rays per taxel and keeps the nearest Moeller-Trumbore hit; the
// hit test depends only on (ray, triangle), so instead of rebuilding a BVH every update we
//   (a) tactile_bin:    splat every contact-surface triangle into the taxels its footprint (in the
//                       sensor frame) can touch — per-taxel atomics on the bin counters; count pass,
//                       exclusive scan, fill pass, so bins are sized exactly;
//   (b) tactile_raster: one warp per taxel walks its bin with the SAME float32 Moeller-Trumbore
//                       arithmetic (bvh.cpp:49-74) for each of its S*S sample rays, stores the weighted
//                       sample pressures in a shared-memory tile and sums them in the reference's
//                       (i, j) order, so the taxel value is bit-identical whenever the nearest hit is
//                       unique (ties between coplanar neighbours carry the same pressure).
#include <algorithm>

#include "dmath.cuh"
#include "hcs_internal.h"

namespace hcs {

struct F3 {
	float x, y, z;
};
__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{ x, y, z }; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 crossf(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float dotf(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // float3.h order

// FILL == false: count the (triangle, taxel) overlaps per taxel; FILL == true: write the triangle ids into
// the exactly-sized bins delimited by the exclusive scan of the counts.
// One pass over the triangle pool serves every sensor (a scene with five pads would otherwise read the whole
// pool five times per pass); grid-stride over the triangles actually emitted this step.
template <bool FILL>
__device__ __forceinline__ void bin_triangle(const SensorDev &sd, const StepIO &io, const TactileTri &t, int i)
{
	int env          = t.env;
	const double *R  = io.xmat + ((size_t)env * io.n_geoms + sd.geom) * 9;
	const double *xp = io.xpos + ((size_t)env * io.n_geoms + sd.geom) * 3;
	double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		double d[3] = { (double)t.v[3 * k] - xp[0], (double)t.v[3 * k + 1] - xp[1], (double)t.v[3 * k + 2] - xp[2] };
		double l[3] = { R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
			            R[2] * d[0] + R[5] * d[1] + R[8] * d[2] };
#pragma unroll
		for (int a = 0; a < 3; ++a)
			lo[a] = fmin(lo[a], l[a]), hi[a] = fmax(hi[a], l[a]);
	}
	double res = sd.resolution, m = 0.01 * res;
	double zs  = sd.size[2];
	if (hi[2] < -m || lo[2] > 1.5 * zs + m)
		return; // outside the accepted ray interval 0 < t < 1.5 zs
	int ix0 = (int)floor((lo[0] - m + sd.size[0]) / res), ix1 = (int)floor((hi[0] + m + sd.size[0]) / res);
	int iy0 = (int)floor((lo[1] - m + sd.size[1]) / res), iy1 = (int)floor((hi[1] + m + sd.size[1]) / res);
	ix0 = max(ix0, 0), iy0 = max(iy0, 0), ix1 = min(ix1, sd.cx - 1), iy1 = min(iy1, sd.cy - 1);
	int ntax = sd.cx * sd.cy;
	for (int x = ix0; x <= ix1; ++x)
		for (int y = iy0; y <= iy1; ++y) {
			int cell = env * ntax + x + sd.cx * y; // bins are indexed x + cx*y (unique)
			if (!FILL) {
				atomicAdd(sd.bin_count + cell, 1);
			} else {
				int slot = sd.bin_offset[cell] + atomicAdd(sd.bin_cursor + cell, 1);
				if (slot < sd.items_cap)
					sd.bin_items[slot] = i;
				else
					atomicOr(io.flags, 4);
			}
		}
}

template <bool FILL>
__global__ void __launch_bounds__(256) tactile_bin_kernel(const SensorDev *sensors, int n_sensors, StepIO io,
                                                          const PairDesc *pairs)
{
	const int n = min(*io.tri_count, io.max_tris);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const TactileTri &t = io.tri_pool[i];
		const PairDesc &P   = pairs[t.key_hi >> TRI_PAIR_SHIFT];
		for (int k = 0; k < n_sensors; ++k)
			if (P.gM == sensors[k].geom || P.gN == sensors[k].geom)
				bin_triangle<FILL>(sensors[k], io, t, i);
	}
}

// ---- exclusive scan of the per-taxel counts (3 phases, 1024-element tiles) ----------------------------------
constexpr int SCAN_TILE = 1024;
__global__ void __launch_bounds__(256) scan_tiles_kernel(const int32_t *in, int32_t *out, int32_t *tile_sums, int n)
{
	__shared__ int warp_sums[8];
	int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
	int v[4], sum = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		v[k] = base + k < n ? in[base + k] : 0;
		sum += v[k];
	}
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	int incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o)
			incl += t;
	}
	if (lane == 31)
		warp_sums[w] = incl;
	__syncthreads();
	int woff = 0;
	for (int k = 0; k < w; ++k)
		woff += warp_sums[k];
	int excl = woff + incl - sum;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if (base + k < n)
			out[base + k] = excl;
		excl += v[k];
	}
	if (threadIdx.x == 255)
		tile_sums[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t *tile_sums, int n_tiles, int32_t *total_out)
{
	__shared__ int warp_sums[32];
	__shared__ int carry;
	if (threadIdx.x == 0)
		carry = 0;
	__syncthreads();
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int base = 0; base < n_tiles; base += 1024) {
		int i    = base + threadIdx.x;
		int v    = i < n_tiles ? tile_sums[i] : 0;
		int incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o)
				incl += t;
		}
		if (lane == 31)
			warp_sums[w] = incl;
		__syncthreads();
		int woff = 0;
		for (int k = 0; k < w; ++k)
			woff += warp_sums[k];
		int c = carry;
		if (i < n_tiles)
			tile_sums[i] = c + woff + incl - v;
		__syncthreads();
		if (threadIdx.x == 1023)
			carry = c + woff + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0)
		*total_out = carry;
}
__global__ void __launch_bounds__(256) scan_add_kernel(int32_t *out, const int32_t *tile_sums, int n)
{
	int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
	int off  = tile_sums[blockIdx.x];
#pragma unroll
	for (int k = 0; k < 4; ++k)
		if (base + k < n)
			out[base + k] += off;
}

// n <= SCAN_TILE (one environment with a 16 x 16 or 32 x 32 array: the reference's own case): the whole scan in one launch
__global__ void __launch_bounds__(256) scan_single_kernel(const int32_t *in, int32_t *out, int n, int32_t *total_out)
{
	__shared__ int warp_sums[8];
	const int base = threadIdx.x * 4;
	int v[4], sum = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		v[k] = base + k < n ? in[base + k] : 0;
		sum += v[k];
	}
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	int incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o)
			incl += t;
	}
	if (lane == 31)
		warp_sums[w] = incl;
	__syncthreads();
	int woff = 0;
	for (int k = 0; k < w; ++k)
		woff += warp_sums[k];
	int excl = woff + incl - sum;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if (base + k < n)
			out[base + k] = excl;
		excl += v[k];
	}
	if (threadIdx.x == 255)
		*total_out = woff + incl;
}

void launch_exclusive_scan_i32(const int32_t *in, int32_t *out, int32_t *tile_tmp, int n, int32_t *total, cudaStream_t s)
{
	if (n <= 0)
		return;
	if (n <= SCAN_TILE) {
		scan_single_kernel<<<1, 256, 0, s>>>(in, out, n, total);
		return;
	}
	const int n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	scan_tiles_kernel<<<n_tiles, 256, 0, s>>>(in, out, tile_tmp, n);
	scan_sums_kernel<<<1, 1024, 0, s>>>(tile_tmp, n_tiles, total);
	scan_add_kernel<<<n_tiles, 256, 0, s>>>(out, tile_tmp, n);
}

constexpr int RASTER_WARPS = 4;
// taxels a persistent warp takes at a time (HCS_RASTER_GROUP overrides).  Lit taxels come in clusters (the contact patch), so
// coarse groups leave a tail of heavy ones: tactile stage of C2 box x 1024 envs 0.925 / 0.633 / 0.579 / 0.560 / 0.579 ms for
// 32 / 8 / 4 / 2 / 1, C5 x 1024 17.6 / 14.3 / 13.9 / 13.4 / 13.3 ms (before the persistent warps: 0.994 and 16.0 ms)
constexpr int RASTER_GROUP = 2;
constexpr int RANK_MAX     = 128; // bins up to this size get an order-independent tie-break
// Triangles of a bin whose sample footprints are staged in shared memory at a time: 64 where the sample arrays are large
// (S = 20: 8 KB per warp, the stage decides how many CTAs fit), 128 where they are small and the bins deep (C5: S = 8,
// hundreds of triangles per taxel: fewer passes of the per-chunk scan).
__host__ __device__ inline int raster_chunk(int S) { return S <= 12 ? 128 : 64; }
// per warp, behind the sample arrays: fp_box int4[chunk] | fp_xy float[6 chunk] | fp_cnt int[chunk + 4] | rank_k u8[RANK_MAX]
__host__ __device__ inline size_t raster_extra_bytes(int S)
{
	const size_t c = (size_t)raster_chunk(S);
	return c * 16 + c * 24 + (c + 4) * 4 + RANK_MAX; // 64: 2960 B (>= the 1536 B of the phase-0 products at S = 32), 128: 5776 B
}
__host__ __device__ inline size_t raster_warp_bytes(int S)
{
	return (((size_t)S * S * 20 + 15) & ~(size_t)15) + raster_extra_bytes(S); // keys + origins, rounded to 16 B, + extras
}

// Moeller-Trumbore, bvh.cpp:49-74, float32 and in the reference's operation order
__device__ __forceinline__ bool moller_trumbore(F3 O, F3 D, const TactileTri &t, float &tt, float &u, float &v)
{
	F3 v0 = f3(t.v[0], t.v[1], t.v[2]), v1 = f3(t.v[3], t.v[4], t.v[5]), v2 = f3(t.v[6], t.v[7], t.v[8]);
	F3 edge1 = v1 - v0, edge2 = v2 - v0;
	F3 h    = crossf(D, edge2);
	float a = dotf(edge1, h);
	if (fabsf(a) < 1e-10f)
		return false;
	float f = 1.0f / a;
	F3 sv   = O - v0;
	u       = f * dotf(sv, h);
	if (u < 0.0f || u > 1.0f)
		return false;
	F3 q = crossf(sv, edge1);
	v    = f * dotf(D, q);
	if (v < 0.0f || u + v > 1.0f)
		return false;
	tt = f * dotf(edge2, q);
	return tt > 0.0f;
}

// One warp per lit taxel.  Dynamic shared memory per warp: S*S 64-bit depth keys, S*S float3 ray origins,
// per-chunk triangle footprints/offsets and bin ranks.
//   phase 0  lanes over samples: ray origins exactly as flat_tactile_sensor.cpp:324-337, keys = +inf
//   phase 1  lanes over the taxel's binned triangles: each triangle is splatted onto the samples inside its
//            footprint (sensor frame), every hit does atomicMin(key[sample], t << 32 | tie-break) in shared
//            memory -> nearest hit per sample, ties resolved by a (pair, order) rank, never by arrival order
//   phase 2  lanes over samples: winner's barycentric pressure * window weight (float/double mix of :346-391)
//   phase 3  lane 0 sums the samples in the reference's (i, j) order -> bit-identical float accumulation
// Round 2: persistent warps.  The grid is what fits the SMs; a warp pulls small groups of consecutive taxels from a counter,
// its lanes look at one taxel each (empty ones get their 0 right there, one coalesced store), the lit ones are then
// rasterised one after the other.  Before, every taxel had its own warp and every four their CTA, whose 57 KB of shared
// memory stayed allocated until its slowest warp was done while the warps of empty taxels had long left (ncu, C2 box x
// 1024 envs: 15 % of the warp slots active).  Footprints are staged 64 triangles at a time and the rank table is bytes:
// 11 KB per warp at S = 20, 5 CTAs per SM instead of 3.
__device__ __forceinline__ void raster_taxel(const SensorDev &sd, const StepIO &io, long unit, int env, int x, int y, float *out,
                                             unsigned char *wsmem, int lane)
{
	int first = sd.bin_offset[unit];
	int n     = min(sd.bin_count[unit], max(sd.items_cap - first, 0));
	if (n <= 0) {
		if (lane == 0)
			*out = 0.0f;
		return;
	}
	const int S = sd.S, S2 = S * S;
	size_t per_warp = raster_warp_bytes(S);
	unsigned long long *key = reinterpret_cast<unsigned long long *>(wsmem);
	float *org              = reinterpret_cast<float *>(key + S2); // [3][S2]
	// per triangle of the chunk: i0, j0, j1, tie (16-byte aligned: starts at the rounded-up array size)
	const int FP_CHUNK      = raster_chunk(S);
	int4 *fp_box            = reinterpret_cast<int4 *>(wsmem + (per_warp - raster_extra_bytes(S)));
	float *fp_xy            = reinterpret_cast<float *>(fp_box + FP_CHUNK);  // 2-D triangle vertices, taxel frame
	int *fp_cnt             = reinterpret_cast<int *>(fp_xy + 6 * FP_CHUNK); // footprint sizes -> exclusive offsets
	unsigned char *rank_k   = reinterpret_cast<unsigned char *>(fp_cnt + FP_CHUNK + 4); // rank -> position in the bin
	const int32_t *items = sd.bin_items + first;
	const double *R  = io.xmat + ((size_t)env * io.n_geoms + sd.geom) * 9;
	const double *xp = io.xpos + ((size_t)env * io.n_geoms + sd.geom) * 3;
	double rot[9];
#pragma unroll
	for (int k = 0; k < 9; ++k)
		rot[k] = R[k];
	float xs = (float)sd.size[0], ys = (float)sd.size[1], zs = (float)sd.size[2];
	F3 normal         = f3((float)rot[2], (float)rot[5], (float)rot[8]);
	double topleft[3] = { (double)(-xs), (double)(-ys), (double)zs };
	double resolution = sd.resolution;
	float rS = sd.rS, rmean = sd.rmean;
	F3 D   = f3(-normal.x, -normal.y, -normal.z);
	F3 off = normal * (float)1e-8;
	double tmax = 1.5 * zs;
	// ---- phase 0
	// pos[0] depends on i only, pos[1] on j only: tabulate the products rot[r][0]*pos0(i) and rot[r][1]*pos1(j)
	// (the sums below are evaluated in the reference's order, so every rounding is unchanged)
	double *prod = reinterpret_cast<double *>(fp_box); // [2][3][S] doubles in the (still unused) phase-1 scratch
	if (lane < S) {
		double p0 = topleft[0] + x * resolution + (double)((float)lane * rS) + 0.5 * (double)rS;
		double p1 = topleft[1] + y * resolution + (double)((float)lane * rS) + 0.5 * (double)rS;
#pragma unroll
		for (int r = 0; r < 3; ++r) {
			prod[r * S + lane]       = rot[3 * r] * p0;
			prod[(3 + r) * S + lane] = rot[3 * r + 1] * p1;
		}
	}
	__syncwarp();
	double pz   = 1.5 * (double)zs;
	double c[3] = { rot[2] * pz, rot[5] * pz, rot[8] * pz };
	int i = lane / S, j = lane - i * S; // (i, j) of sample s = i * S + j, advanced by 32 samples per iteration without a division
	const int di = 32 / S, dj = 32 - di * S;
	for (int s = lane; s < S2; s += 32) {
		double w[3] = { prod[i] + prod[3 * S + j] + c[0], prod[S + i] + prod[4 * S + j] + c[1], prod[2 * S + i] + prod[5 * S + j] + c[2] };
		w[0] += xp[0], w[1] += xp[1], w[2] += xp[2];
		F3 O = f3((float)w[0], (float)w[1], (float)w[2]) + off;
		org[s] = O.x, org[S2 + s] = O.y, org[2 * S2 + s] = O.z;
		key[s] = ~0ull;
		i += di, j += dj;
		if (j >= S)
			j -= S, ++i;
	}
	__syncwarp();
	// ---- phase 1: splat.  Triangles are taken in chunks of CHUNK; lanes first compute each triangle's sample
	// footprint and tie-break rank, a warp scan turns the footprint sizes into item offsets, then the lanes
	// share the flattened (triangle, sample-row) items evenly; each item tests only the samples of its row that
	// the triangle can cover (fan triangles are slivers: their boxes are mostly empty).
	bool ranked = n <= RANK_MAX;
	double base_x = topleft[0] + x * resolution, base_y = topleft[1] + y * resolution;
	const double margin = 1e-5; // >> float32 rounding of the hit test at these magnitudes, << sample spacing
	for (int c0 = 0; c0 < n; c0 += FP_CHUNK) {
		int m = min(FP_CHUNK, n - c0);
		for (int k = lane; k < m; k += 32) {
			const TactileTri &t = io.tri_pool[items[c0 + k]];
			double lo[2] = { 1e300, 1e300 }, hi[2] = { -1e300, -1e300 }, lxy[3][2];
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				double d[3] = { (double)t.v[3 * c] - xp[0], (double)t.v[3 * c + 1] - xp[1], (double)t.v[3 * c + 2] - xp[2] };
				double lx = rot[0] * d[0] + rot[3] * d[1] + rot[6] * d[2], ly = rot[1] * d[0] + rot[4] * d[1] + rot[7] * d[2];
				lxy[c][0] = lx, lxy[c][1] = ly;
				lo[0] = fmin(lo[0], lx), hi[0] = fmax(hi[0], lx), lo[1] = fmin(lo[1], ly), hi[1] = fmax(hi[1], ly);
			}
			// sample i sits at base + (i + 0.5) * rS
			int i0 = max(0, (int)ceil((lo[0] - margin - base_x) / (double)rS - 0.5)),
			    i1 = min(S - 1, (int)floor((hi[0] + margin - base_x) / (double)rS - 0.5));
			int j0 = max(0, (int)ceil((lo[1] - margin - base_y) / (double)rS - 0.5)),
			    j1 = min(S - 1, (int)floor((hi[1] + margin - base_y) / (double)rS - 0.5));
			int rows = max(0, i1 - i0 + 1), cols = max(0, j1 - j0 + 1);
			int r = c0 + k;
			if (ranked) { // order-independent tie-break: rank under the (pair, order) key
				r = 0;
				for (int k2 = 0; k2 < n; ++k2) {
					const TactileTri &o = io.tri_pool[items[k2]];
					r += (o.key_hi < t.key_hi) ||
					     (o.key_hi == t.key_hi && (o.key_lo < t.key_lo || (o.key_lo == t.key_lo && items[k2] < items[c0 + k])));
				}
				rank_k[r] = (unsigned char)(c0 + k);
			}
			fp_cnt[k] = cols > 0 ? rows : 0; // one item per sample row of the footprint
			fp_box[k] = make_int4(i0, j0, j1, r);
#pragma unroll
			for (int c = 0; c < 3; ++c) { // 2-D vertices in the sensor frame, relative to the taxel corner
				fp_xy[6 * k + 2 * c]     = (float)(lxy[c][0] - base_x);
				fp_xy[6 * k + 2 * c + 1] = (float)(lxy[c][1] - base_y);
			}
		}
		__syncwarp();
		// exclusive scan of the (at most FP_CHUNK) footprint sizes, entry m = their total: 4 consecutive entries per lane
		int v4[4], sum = 0;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			int idx = 4 * lane + q;
			v4[q]   = idx < m ? fp_cnt[idx] : 0;
			sum += v4[q];
		}
		int incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o)
				incl += t;
		}
		int total = __shfl_sync(0xffffffffu, incl, 31);
		int excl  = incl - sum;
		__syncwarp();
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			int idx = 4 * lane + q;
			if (idx <= m)
				fp_cnt[idx] = excl;
			excl += v4[q];
		}
		__syncwarp();
		for (int w = lane; w < total; w += 32) {
			int lo_k = 0, hi_k = m; // largest k with offset[k] <= w
			while (hi_k - lo_k > 1) {
				int mid = (lo_k + hi_k) >> 1;
				if (fp_cnt[mid] <= w)
					lo_k = mid;
				else
					hi_k = mid;
			}
			int4 b = fp_box[lo_k];
			int i  = b.x + (w - fp_cnt[lo_k]);
			// conservative column span of row i: y-extent of (triangle  ∩  strip |x - x_i| <= m), widened by m.
			// Any sample within the float32 hit tolerance (~1e-6 m) of the triangle lies inside this span.
			const float m2 = 2e-5f;
			float xi = ((float)i + 0.5f) * rS, xa = xi - m2, xb = xi + m2;
			float ymin = 1e30f, ymax = -1e30f;
			const float *q = fp_xy + 6 * lo_k;
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				float px = q[2 * c], py = q[2 * c + 1], qx = q[2 * ((c + 1) % 3)], qy = q[2 * ((c + 1) % 3) + 1];
				if (px >= xa && px <= xb)
					ymin = fminf(ymin, py), ymax = fmaxf(ymax, py);
				float dx = qx - px;
				if ((px - xa) * (qx - xa) < 0.f) {
					float yy = py + (xa - px) / dx * (qy - py);
					ymin = fminf(ymin, yy), ymax = fmaxf(ymax, yy);
				}
				if ((px - xb) * (qx - xb) < 0.f) {
					float yy = py + (xb - px) / dx * (qy - py);
					ymin = fminf(ymin, yy), ymax = fmaxf(ymax, yy);
				}
			}
			int ja = max(b.y, (int)ceilf((ymin - m2) / rS - 0.5f)), jb = min(b.z, (int)floorf((ymax + m2) / rS - 0.5f));
			const TactileTri &t = io.tri_pool[items[c0 + lo_k]];
			// moller_trumbore() with everything that depends only on (D, triangle) hoisted out of the row: all rays
			// of a sensor are parallel, so h, a and f are computed once per item (same operations, same rounding)
			F3 v0 = f3(t.v[0], t.v[1], t.v[2]), v1 = f3(t.v[3], t.v[4], t.v[5]), v2 = f3(t.v[6], t.v[7], t.v[8]);
			F3 edge1 = v1 - v0, edge2 = v2 - v0;
			F3 h     = crossf(D, edge2);
			float a  = dotf(edge1, h);
			if (fabsf(a) < 1e-10f)
				continue;
			float f = 1.0f / a;
			for (int j = ja; j <= jb; ++j) {
				int s   = i * S + j;
				F3 sv   = f3(org[s], org[S2 + s], org[2 * S2 + s]) - v0;
				float u = f * dotf(sv, h);
				if (u < 0.0f || u > 1.0f)
					continue;
				F3 q    = crossf(sv, edge1);
				float v = f * dotf(D, q);
				if (v < 0.0f || u + v > 1.0f)
					continue;
				float tt = f * dotf(edge2, q);
				if (tt > 0.0f)
					atomicMin(&key[s], ((unsigned long long)__float_as_uint(tt) << 32) | (unsigned)b.w);
			}
		}
		__syncwarp();
	}
	// ---- phase 2: winners -> weighted sample pressures, written over the x plane of the origins (sample s is read and
	// written by the same lane), so that phase 3 reads four values per 128-bit load
	float *val = org;
	for (int s = lane; s < S2; s += 32) {
		unsigned long long kk = key[s];
		float value = 0.0f;
		if (kk != ~0ull) {
			unsigned tie = (unsigned)(kk & 0xffffffffu);
			int k        = ranked ? rank_k[tie] : (int)tie;
			const TactileTri &t = io.tri_pool[items[k]];
			F3 O = f3(org[s], org[S2 + s], org[2 * S2 + s]);
			float tt, u, v;
			if (moller_trumbore(O, D, t, tt, u, v) && (double)tt < tmax) {
				double b0 = (double)(1 - u - v), b1 = (double)u, b2 = (double)v;
				double ev = b0 * t.e[0];
				ev += b1 * t.e[1];
				ev += b2 * t.e[2];
				float raw = (float)(ev * (double)rmean);
				value     = sd.weights[s] * raw;
			}
		}
		val[s] = value;
	}
	__syncwarp();
	// ---- phase 3
	if (lane == 0) {
		float avg = 0;
		const float4 *v4 = reinterpret_cast<const float4 *>(val); // (org starts 16-byte aligned: S2 keys of 8 bytes behind a 16-byte base)
		int s = 0;
		if ((S2 & 1) == 0) { // key array = 8 * S2 bytes: org is 16-byte aligned when S2 is even
#pragma unroll 4
			for (; s + 4 <= S2; s += 4) { // loads are independent (batched by the unroll); the adds stay in order
				const float4 q = v4[s >> 2];
				avg += q.x;
				avg += q.y;
				avg += q.z;
				avg += q.w;
			}
		}
		for (; s < S2; ++s)
			avg += val[s];
		*out = avg;
	}
	__syncwarp(); // the warp's next taxel reuses the arrays
}

__global__ void __launch_bounds__(32 * RASTER_WARPS) tactile_raster_kernel(SensorDev sd, StepIO io, int group)
{
	extern __shared__ __align__(16) unsigned char raster_smem[];
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char *wsmem = raster_smem + wib * raster_warp_bytes(sd.S);
	const int ntax   = sd.cx * sd.cy;
	const long total = (long)io.n_env * ntax;
	for (;;) {
		int g = 0;
		if (lane == 0)
			g = atomicAdd(sd.raster_counter, 1);
		g = __shfl_sync(0xffffffffu, g, 0);
		const long u0 = (long)g * group;
		if (u0 >= total)
			break;
		// lane l < RASTER_GROUP looks at taxel u0 + l: empty ones are done here
		const long unit = u0 + lane;
		bool lit = false;
		if (lane < group && unit < total) {
			const int env = (int)(unit / ntax), cell = (int)(unit - (long)env * ntax);
			const int x = cell % sd.cx, y = cell / sd.cx;
			// the reference stores taxel (x,y) at x + cy*y (flat_tactile_sensor.cpp:396-397; quirk Q8: only a
			// bijection when cx == cy); out-of-range indices of non-square arrays are dropped
			const int oidx = x + sd.cy * y;
			if (oidx < ntax) {
				lit = sd.bin_count[unit] > 0;
				if (!lit)
					sd.image[(size_t)env * ntax + oidx] = 0.0f;
			}
		}
		unsigned m = __ballot_sync(0xffffffffu, lit);
		while (m) {
			const int src = __ffs(m) - 1;
			m &= m - 1;
			const long u  = u0 + src;
			const int env = (int)(u / ntax), cell = (int)(u - (long)env * ntax);
			const int x = cell % sd.cx, y = cell / sd.cx;
			raster_taxel(sd, io, u, env, x, y, sd.image + (size_t)env * ntax + (x + sd.cy * y), wsmem, lane);
		}
	}
}

__global__ void tactile_clear_kernel(SensorDev sd, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		sd.bin_count[i]  = 0;
		sd.bin_cursor[i] = 0;
	}
	if (i == 0)
		*sd.raster_counter = 0;
}

// per sensor: clear, 3-phase scan, raster; for all sensors together: one count pass and one fill pass over the pool
int launch_tactile(const SensorDev *sensors, const SensorDev *d_sensors, int n_sensors, const StepIO &io,
                   const PairDesc *d_pairs, cudaStream_t s)
{
	if (n_sensors <= 0)
		return 0;
	int launches = 0;
	int tgrid    = (int)std::min<long>(((long)io.max_tris + 255) / 256, (long)io.n_sms * 8);
	for (int k = 0; k < n_sensors; ++k) {
		int ncell = io.n_env * sensors[k].cx * sensors[k].cy;
		tactile_clear_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(sensors[k], ncell);
		++launches;
	}
	if (io.max_tris > 0) {
		tactile_bin_kernel<false><<<tgrid, 256, 0, s>>>(d_sensors, n_sensors, io, d_pairs);
		++launches;
	}
	for (int k = 0; k < n_sensors; ++k) {
		const SensorDev &sd = sensors[k];
		int ncell = io.n_env * sd.cx * sd.cy, n_tiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
		launch_exclusive_scan_i32(sd.bin_count, sd.bin_offset, sd.scan_tmp, ncell, sd.bin_offset + ncell, s);
		launches += ncell <= SCAN_TILE ? 1 : 3;
	}
	if (io.max_tris > 0) {
		tactile_bin_kernel<true><<<tgrid, 256, 0, s>>>(d_sensors, n_sensors, io, d_pairs);
		++launches;
	}
	for (int k = 0; k < n_sensors; ++k) {
		const SensorDev &sd = sensors[k];
		long ncell      = (long)io.n_env * sd.cx * sd.cy;
		int raster_smem = RASTER_WARPS * (int)raster_warp_bytes(sd.S);
		ensure_dynamic_smem(tactile_raster_kernel, raster_smem);
		// persistent warps: as many CTAs as the SMs hold (shared memory; 80 registers allow 6), never more than groups
		const int per_sm = std::max(1, std::min(6, (int)((227 * 1024) / (raster_smem + 1024))));
		static const int group = getenv("HCS_RASTER_GROUP") ? std::max(1, std::min(32, atoi(getenv("HCS_RASTER_GROUP")))) : RASTER_GROUP;
		const long groups = (ncell + group - 1) / group;
		const int grid    = (int)std::max<long>(1, std::min<long>((long)io.n_sms * per_sm, (groups + RASTER_WARPS - 1) / RASTER_WARPS));
		tactile_raster_kernel<<<grid, 32 * RASTER_WARPS, raster_smem, s>>>(sd, io, group);
		++launches;
	}
	return launches;
}

// =====================================================================================================
// K9 curved sensor (CurvedSensor::internal_update, SENS/src/curved_sensor.cpp:388-481).  The reference rebuilds a
// BLAS/TLAS over the contact surfaces every update and casts one ray per (taxel, assigned surface sample); here
// every distinct sample casts once: triangles are binned into the static ray grid (count, scan, fill), one warp
// per (env, cell) finds the nearest hit of each of the cell's rays in the cell's bin with the reference's float32
// Moeller-Trumbore (ties: smallest (pair, slice, index, fan triangle) key), then one thread per taxel adds
// weight * pressure over its sample list in the reference's order.
// =====================================================================================================
__global__ void curved_clear_kernel(CurvedDev cd, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		cd.bin_count[i]  = 0;
		cd.bin_cursor[i] = 0;
	}
}

template <bool FILL>
__global__ void __launch_bounds__(256) curved_bin_kernel(CurvedDev cd, StepIO io, const PairDesc *pairs)
{
	const int n = min(*io.tri_count, io.max_tris);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const TactileTri &t = io.tri_pool[i];
		const PairDesc &P   = pairs[t.key_hi >> TRI_PAIR_SHIFT];
		if (P.gM != cd.geom && P.gN != cd.geom)
			continue;
		const int env    = t.env;
		const double *R  = io.xmat + ((size_t)env * io.n_geoms + cd.geom) * 9;
		const double *xp = io.xpos + ((size_t)env * io.n_geoms + cd.geom) * 3;
		double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			double d[3] = { (double)t.v[3 * k] - xp[0], (double)t.v[3 * k + 1] - xp[1], (double)t.v[3 * k + 2] - xp[2] };
			double l[3] = { R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
				            R[2] * d[0] + R[5] * d[1] + R[8] * d[2] };
#pragma unroll
			for (int a = 0; a < 3; ++a)
				lo[a] = fmin(lo[a], l[a]), hi[a] = fmax(hi[a], l[a]);
		}
		// a ray leaves its cell by at most its length: grow the triangle's box by that (+ float32 slack)
		const double m = cd.include_margin * 1.001 + 1e-6;
		int c0[3], c1[3];
		bool out = false;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			c0[a] = max((int)floor((lo[a] - m - cd.origin[a]) / cd.cell), 0);
			c1[a] = min((int)floor((hi[a] + m - cd.origin[a]) / cd.cell), cd.dims[a] - 1);
			out |= c0[a] > c1[a];
		}
		if (out)
			continue;
		for (int x = c0[0]; x <= c1[0]; ++x)
			for (int y = c0[1]; y <= c1[1]; ++y)
				for (int z = c0[2]; z <= c1[2]; ++z) {
					int id = cd.cell_lookup[((size_t)x * cd.dims[1] + y) * cd.dims[2] + z];
					if (id < 0)
						continue;
					int cell = env * cd.n_cells + id;
					if (!FILL) {
						atomicAdd(cd.bin_count + cell, 1);
					} else {
						int slot = cd.bin_offset[cell] + atomicAdd(cd.bin_cursor + cell, 1);
						if (slot < cd.items_cap)
							cd.bin_items[slot] = i;
						else
							atomicOr(io.flags, 4);
					}
				}
	}
}

__global__ void __launch_bounds__(128) curved_cast_kernel(CurvedDev cd, StepIO io)
{
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const long unit = (long)blockIdx.x * 4 + wib;
	if (unit >= (long)io.n_env * cd.n_cells)
		return;
	const int env = (int)(unit / cd.n_cells), cell = (int)(unit - (long)env * cd.n_cells);
	const int first = cd.bin_offset[unit];
	const int n     = min(cd.bin_count[unit], max(cd.items_cap - first, 0));
	const int32_t *items = cd.bin_items + first;
	const double *R  = io.xmat + ((size_t)env * io.n_geoms + cd.geom) * 9;
	const double *xp = io.xpos + ((size_t)env * io.n_geoms + cd.geom) * 3;
	for (int k = cd.cell_ray_off[cell] + lane; k < cd.cell_ray_off[cell + 1]; k += 32) {
		const int ray    = cd.cell_rays[k];
		const double *p  = cd.ray_pos + 3 * (size_t)ray, *nn = cd.ray_nrm + 3 * (size_t)ray;
		double w[3], wn[3]; // M * (p, 1), M * (n, 0) with M = [R | x], curved_sensor.cpp:391-399
#pragma unroll
		for (int r = 0; r < 3; ++r) {
			w[r]  = R[3 * r] * p[0] + R[3 * r + 1] * p[1] + R[3 * r + 2] * p[2] + xp[r];
			wn[r] = R[3 * r] * nn[0] + R[3 * r + 1] * nn[1] + R[3 * r + 2] * nn[2];
		}
		F3 normal = f3((float)wn[0], (float)wn[1], (float)wn[2]);
		F3 O      = f3((float)w[0], (float)w[1], (float)w[2]) + normal * (float)1e-8;
		F3 D      = f3(-normal.x, -normal.y, -normal.z);
		float best_t = 1e30f, best_u = 0, best_v = 0;
		unsigned best_ps = 0xffffffffu, best_idx = 0xffffffffu;
		int best = -1;
		for (int j = 0; j < n; ++j) { // every lane walks the same bin: the loads are broadcasts
			const int item      = items[j];
			const TactileTri &t = io.tri_pool[item];
			float tt, u, v;
			if (moller_trumbore(O, D, t, tt, u, v)) {
				bool better = tt < best_t ||
				              (tt == best_t && (t.key_hi < best_ps || (t.key_hi == best_ps && t.key_lo < best_idx)));
				if (better)
					best_t = tt, best_u = u, best_v = v, best_ps = t.key_hi, best_idx = t.key_lo, best = item;
			}
		}
		double raw = 0;
		if (best >= 0 && (double)best_t < cd.include_margin) { // t > 0 is part of the hit test
			const TactileTri &t = io.tri_pool[best];
			double b0 = (double)(1 - best_u - best_v), b1 = (double)best_u, b2 = (double)best_v;
			raw = b0 * t.e[0];
			raw += b1 * t.e[1];
			raw += b2 * t.e[2];
		}
		cd.raw[(size_t)env * cd.n_rays + ray] = raw;
	}
}

__global__ void __launch_bounds__(128) curved_taxel_kernel(CurvedDev cd, StepIO io)
{
	long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long)io.n_env * cd.n_taxels)
		return;
	int env = (int)(i / cd.n_taxels), taxel = (int)(i - (long)env * cd.n_taxels);
	const double *raw = cd.raw + (size_t)env * cd.n_rays;
	double pressure   = 0;
	for (int k = cd.taxel_off[taxel]; k < cd.taxel_off[taxel + 1]; ++k) {
		double r = raw[cd.taxel_ray[k]];
		if (r != 0.0) // the reference adds only accepted hits
			pressure += cd.taxel_w[k] * r;
	}
	cd.values[i] = (float)pressure;
}

int launch_curved(const CurvedDev &cd, const StepIO &io, const PairDesc *d_pairs, cudaStream_t s)
{
	const int ncell = io.n_env * cd.n_cells;
	if (ncell <= 0 || cd.n_taxels <= 0)
		return 0;
	const int n_tiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
	const int tgrid   = (int)std::max<long>(1, std::min<long>(((long)io.max_tris + 255) / 256, (long)io.n_sms * 8));
	curved_clear_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(cd, ncell);
	curved_bin_kernel<false><<<tgrid, 256, 0, s>>>(cd, io, d_pairs);
	scan_tiles_kernel<<<n_tiles, 256, 0, s>>>(cd.bin_count, cd.bin_offset, cd.scan_tmp, ncell);
	scan_sums_kernel<<<1, 1024, 0, s>>>(cd.scan_tmp, n_tiles, cd.bin_offset + ncell);
	scan_add_kernel<<<n_tiles, 256, 0, s>>>(cd.bin_offset, cd.scan_tmp, ncell);
	curved_bin_kernel<true><<<tgrid, 256, 0, s>>>(cd, io, d_pairs);
	curved_cast_kernel<<<(ncell + 3) / 4, 128, 0, s>>>(cd, io);
	long nt = (long)io.n_env * cd.n_taxels;
	curved_taxel_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(cd, io);
	return 8;
}

// =====================================================================================================
// K10 taxel sensor (TaxelSensor::internal_update, SENS/src/taxel_sensor.cpp:158-478, sample_method "default").
// The reference samples every contact-surface triangle on a barycentric lattice, builds the dense taxel x sample
// distance matrix and loops over it per taxel.  Here triangles are binned per (env, taxel) by their box grown by
// include_margin; one warp per (env, taxel) walks its bin in a canonical triangle order, each lane regenerating its
// triangles' samples with the reference's double arithmetic.  Quirk Q12 is reproduced: weighted / mean fall
// through to squared, closest keeps the pressure only with visualize, taxels without a sample in range keep their
// previous value, an update without any sample zeroes the message.
// =====================================================================================================
__global__ void taxel_clear_kernel(TaxelDev td, int n_cells, int n_env)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_cells) {
		td.bin_count[i]  = 0;
		td.bin_cursor[i] = 0;
	}
	if (i < n_env)
		td.env_tris[i] = 0;
}

__device__ __forceinline__ void taxel_world(const TaxelDev &td, const double *R, const double *xp, int i, double tw[3])
{
	const double *t = td.taxel_pos + 3 * (size_t)i; // M * (t, 1), taxel_sensor.cpp:273-280
#pragma unroll
	for (int r = 0; r < 3; ++r)
		tw[r] = R[3 * r] * t[0] + R[3 * r + 1] * t[1] + R[3 * r + 2] * t[2] + xp[r];
}

template <bool FILL>
__global__ void __launch_bounds__(256) taxel_bin_kernel(TaxelDev td, StepIO io, const PairDesc *pairs)
{
	const int n = min(*io.tri_count, io.max_tris);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const TactileTri &t = io.tri_pool[i];
		const PairDesc &P   = pairs[t.key_hi >> TRI_PAIR_SHIFT];
		if (P.gM != td.geom && P.gN != td.geom)
			continue;
		const int env = t.env;
		if (!FILL)
			atomicAdd(td.env_tris + env, 1);
		const double *vd = io.tri_vd + 9 * (size_t)i;
		const double m   = td.include_margin * (1 + 1e-9) + 1e-12; // every sample lies inside the triangle's box
		double lo[3], hi[3];
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			lo[a] = fmin(fmin(vd[a], vd[3 + a]), vd[6 + a]) - m;
			hi[a] = fmax(fmax(vd[a], vd[3 + a]), vd[6 + a]) + m;
		}
		const double *R  = io.xmat + ((size_t)env * io.n_geoms + td.geom) * 9;
		const double *xp = io.xpos + ((size_t)env * io.n_geoms + td.geom) * 3;
		for (int k = 0; k < td.n_taxels; ++k) {
			double tw[3];
			taxel_world(td, R, xp, k, tw);
			if (tw[0] < lo[0] || tw[0] > hi[0] || tw[1] < lo[1] || tw[1] > hi[1] || tw[2] < lo[2] || tw[2] > hi[2])
				continue;
			int cell = env * td.n_taxels + k;
			if (!FILL) {
				atomicAdd(td.bin_count + cell, 1);
			} else {
				int slot = td.bin_offset[cell] + atomicAdd(td.bin_cursor + cell, 1);
				if (slot < td.items_cap)
					td.bin_items[slot] = i;
				else
					atomicOr(io.flags, 4);
			}
		}
	}
}

constexpr int TAXEL_RANK_MAX = 256;
__global__ void __launch_bounds__(128) taxel_gather_kernel(TaxelDev td, StepIO io)
{
	__shared__ int order_s[4][TAXEL_RANK_MAX];
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const long unit = (long)blockIdx.x * 4 + wib;
	if (unit >= (long)io.n_env * td.n_taxels)
		return;
	const int env = (int)(unit / td.n_taxels), taxel = (int)(unit - (long)env * td.n_taxels);
	float *out = td.values + unit;
	if (td.env_tris[env] == 0) { // no sample at all: the message is zeroed (:455-477)
		if (lane == 0)
			*out = 0.0f;
		return;
	}
	const int first = td.bin_offset[unit];
	const int n     = min(td.bin_count[unit], max(td.items_cap - first, 0));
	if (n == 0)
		return; // no sample in range: the previous value stays
	const int32_t *items = td.bin_items + first;
	int *order           = order_s[wib];
	const bool ranked    = n <= TAXEL_RANK_MAX;
	if (ranked) { // canonical order of the bin: by the triangles' (pair, slice, index, fan triangle) key
		for (int k = lane; k < n; k += 32) {
			const TactileTri &t = io.tri_pool[items[k]];
			int r = 0;
			for (int k2 = 0; k2 < n; ++k2) {
				const TactileTri &o = io.tri_pool[items[k2]];
				r += (o.key_hi < t.key_hi) || (o.key_hi == t.key_hi && o.key_lo < t.key_lo);
			}
			order[r] = items[k];
		}
		__syncwarp();
	}
	const double *R  = io.xmat + ((size_t)env * io.n_geoms + td.geom) * 9;
	const double *xp = io.xpos + ((size_t)env * io.n_geoms + td.geom) * 3;
	double tw[3];
	taxel_world(td, R, xp, taxel, tw);
	const double tsq = tw[0] * tw[0] + tw[1] * tw[1] + tw[2] * tw[2];
	const double margin = td.include_margin, margin_sq = margin * margin, res = td.sample_resolution;
	double pressure = 0, dmin = 1e300, pmin = 0;
	int ws = 0, kmin = 0x7fffffff;
	for (int k = lane; k < n; k += 32) {
		const int item      = ranked ? order[k] : items[k];
		const TactileTri &t = io.tri_pool[item];
		const double *v     = io.tri_vd + 9 * (size_t)item;
		double e10 = sqrt((v[3] - v[0]) * (v[3] - v[0]) + (v[4] - v[1]) * (v[4] - v[1]) + (v[5] - v[2]) * (v[5] - v[2]));
		double e20 = sqrt((v[6] - v[0]) * (v[6] - v[0]) + (v[7] - v[1]) * (v[7] - v[1]) + (v[8] - v[2]) * (v[8] - v[2]));
		int st0 = (int)(e10 / res) + 1, st1 = (int)(e20 / res) + 1, st2 = st1; // :191-194 (st2 repeats st1)
		int stm = max(st0, st1);
		const double da = 1. / stm, db = 1. / st2;
		for (double a = 0; a <= 1; a += da)
			for (double b = 0; b <= 1; b += db) {
				double b0 = a, b1 = (1 - a) * (1 - b), b2 = (1 - a) * b;
				double p[3];
#pragma unroll
				for (int c = 0; c < 3; ++c)
					p[c] = b0 * v[c] + b1 * v[3 + c] + b2 * v[6 + c];
				double dj = (-2 * (tw[0] * p[0] + tw[1] * p[1] + tw[2] * p[2]) + tsq) + (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
				double pj = b0 * t.e[0] + b1 * t.e[1] + b2 * t.e[2];
				if (td.method == 0) {
					if (dj < dmin) // first minimum in (triangle order, sample order)
						dmin = dj, pmin = pj, kmin = k;
				} else if (dj < margin_sq) {
					double w = fmax(0.0, margin - sqrt(dj));
					pressure += (w * w) * fabs(pj);
					ws += 1;
				}
			}
	}
	if (td.method == 0) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			double d2 = __shfl_xor_sync(0xffffffffu, dmin, o), p2 = __shfl_xor_sync(0xffffffffu, pmin, o);
			int k2 = __shfl_xor_sync(0xffffffffu, kmin, o);
			if (d2 < dmin || (d2 == dmin && k2 < kmin))
				dmin = d2, pmin = p2, kmin = k2;
		}
		if (lane == 0 && dmin < margin_sq)
			*out = (td.visualize && fabs(pmin) > 1e-6) ? (float)pmin : 0.0f; // :312-328
		return;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		pressure += __shfl_xor_sync(0xffffffffu, pressure, o);
		ws += __shfl_xor_sync(0xffffffffu, ws, o);
	}
	if (lane == 0 && ws > 0)
		*out = (float)(pressure * res); // :409-412
}

// ---- sample_method AREA_IMPORTANCE (taxel_sensor.cpp:211-254) ------------------------------------------------------
// The reference walks the triangles of every contact surface of the sensor geom once, spends one stratum of
// sample_resolution * total_area per sample along the cumulative area and draws a uniform barycentric point in the
// triangle that owns the stratum's start, all from ONE std::default_random_engine (minstd_rand0, default seed) that is
// created anew in every update.  Which random numbers a triangle gets depends on the triangle order, and Drake's is
// not observable; ours is canonical: pairs in pair order, polygons by (elemM, elemN), fan triangles in fan order,
// area(t) = |(b - a) x (c - a)| / 2 of the world vertices.  The walk is inherently sequential (one accumulator, one
// random stream): one CTA per environment sorts the environment's triangles in shared memory, thread 0 walks them
// and records (triangle, generator state) per sample, then all threads turn the records into points and pressures
// with libstdc++'s generate_canonical<double, 53> arithmetic (two 31-bit draws per double).
constexpr int TAXEL_AI_CAP = 4096; // triangles per (environment, sensor) the shared-memory sort holds

template <bool FILL>
__global__ void __launch_bounds__(256) taxel_ai_list_kernel(TaxelDev td, StepIO io, const PairDesc *pairs)
{
	const int n = min(*io.tri_count, io.max_tris);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const TactileTri &t = io.tri_pool[i];
		const PairDesc &P   = pairs[t.key_hi >> TRI_PAIR_SHIFT];
		if (P.gM != td.geom && P.gN != td.geom)
			continue;
		if (!FILL)
			atomicAdd(td.env_tris + t.env, 1);
		else
			td.env_items[td.env_offset[t.env] + atomicAdd(td.env_cursor + t.env, 1)] = i;
	}
}

__device__ __forceinline__ unsigned minstd_next(unsigned x) { return (unsigned)((16807ull * x) % 2147483647ull); }
// std::generate_canonical<double, 53>(minstd_rand0) as libstdc++ evaluates it, = uniform_real_distribution(0, 1)
__device__ __forceinline__ double minstd_canonical(unsigned &x)
{
	const double r = 2147483646.0; // max() - min() + 1
	x              = minstd_next(x);
	double sum     = (double)(x - 1u);
	x              = minstd_next(x);
	sum += (double)(x - 1u) * r;
	double ret = sum / (r * r);
	return ret >= 1.0 ? 0.99999999999999988897769753748 : ret; // nextafter(1, 0)
}

__global__ void __launch_bounds__(256) taxel_ai_sample_kernel(TaxelDev td, StepIO io)
{
	extern __shared__ __align__(16) unsigned char ai_smem[];
	unsigned long long *key = reinterpret_cast<unsigned long long *>(ai_smem); // [CAP]
	double *area            = reinterpret_cast<double *>(key + TAXEL_AI_CAP);  // [CAP]
	int *item               = reinterpret_cast<int *>(area + TAXEL_AI_CAP);    // [CAP]
	__shared__ int n_out;
	const int env = blockIdx.x, tid = threadIdx.x;
	const int first = td.env_offset[env];
	int n           = td.env_tris[env];
	if (n > TAXEL_AI_CAP) { // reported, not UB
		if (tid == 0) {
			atomicOr(io.flags, 4);
			td.n_samples[env] = 0;
		}
		return;
	}
	int np2 = 1;
	while (np2 < n)
		np2 <<= 1;
	for (int i = tid; i < np2; i += blockDim.x) {
		if (i < n) {
			const int it        = td.env_items[first + i];
			const TactileTri &t = io.tri_pool[it];
			const uint2 el      = io.tri_elem[it];
			key[i]  = ((unsigned long long)(t.key_hi >> TRI_PAIR_SHIFT) << 55) | ((unsigned long long)el.x << 29) |
			         ((unsigned long long)el.y << 3) | (unsigned long long)(t.key_lo & 7u);
			item[i] = it;
		} else {
			key[i]  = ~0ull;
			item[i] = -1;
		}
	}
	__syncthreads();
	for (int k = 2; k <= np2; k <<= 1) // bitonic sort by key (keys are distinct)
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int i = tid; i < np2; i += blockDim.x) {
				int l = i ^ j;
				if (l > i) {
					bool up = (i & k) == 0;
					unsigned long long a = key[i], b = key[l];
					if ((a > b) == up) {
						key[i] = b, key[l] = a;
						int t = item[i];
						item[i] = item[l], item[l] = t;
					}
				}
			}
			__syncthreads();
		}
	for (int i = tid; i < n; i += blockDim.x) {
		const double *v = io.tri_vd + 9 * (size_t)item[i];
		D3 a = mk(v[0], v[1], v[2]), b = mk(v[3], v[4], v[5]), c = mk(v[6], v[7], v[8]);
		D3 cr   = cross(b - a, c - a);
		area[i] = 0.5 * sqrt(dot(cr, cr));
	}
	__syncthreads();
	int32_t *s_item    = td.env_items + first; // reused: the environment's own slots now hold the sorted order
	double *samp       = td.samples + (size_t)env * td.max_samples * 4;
	unsigned *s_state  = reinterpret_cast<unsigned *>(samp); // records first (8 bytes per sample), points later
	if (tid == 0) {
		unsigned x = 1u; // std::default_random_engine generator;
		int m      = 0;
		bool full  = false;
		for (int i0 = 0; i0 < n && !full;) {
			const unsigned long long pair = key[i0] >> 55;
			int i1 = i0;
			double total = 0;
			while (i1 < n && (key[i1] >> 55) == pair)
				total += area[i1++];
			const double area_resolution = td.sample_resolution * total;
			double acc = 0, at = 0;
			if (area_resolution > 0)
				for (int i = i0; i < i1 && !full; ++i) {
					at += area[i];
					while (acc < at) {
						acc += area_resolution;
						if (m >= td.max_samples) {
							full = true;
							break;
						}
						s_state[2 * m]     = x;
						s_state[2 * m + 1] = (unsigned)i;
						++m;
						x = minstd_next(minstd_next(minstd_next(minstd_next(x)))); // two doubles = four draws
					}
				}
			i0 = i1;
		}
		if (full)
			atomicOr(io.flags, 4);
		n_out             = m;
		td.n_samples[env] = m;
	}
	__syncthreads();
	const int m = n_out;
	// records -> (point, pressure); a record is read before its slot range [4k, 4k + 4) doubles is overwritten only by
	// the thread that owns sample k, and records live in the first m doubles: go from the back in rounds so that no
	// record is overwritten before it is read
	for (int base = ((m - 1) / (int)blockDim.x) * (int)blockDim.x; base >= 0; base -= blockDim.x) {
		const int k = base + tid;
		unsigned x = 0, i = 0;
		if (k < m)
			x = s_state[2 * k], i = s_state[2 * k + 1];
		__syncthreads();
		if (k < m) {
			const int it        = item[i];
			const TactileTri &t = io.tri_pool[it];
			const double *v     = io.tri_vd + 9 * (size_t)it;
			const double u0 = minstd_canonical(x), u1 = minstd_canonical(x);
			const double a = 1.0 - sqrt(u0), b = (1.0 - a) * u1;
			const double b0 = a, b1 = (1 - a) * (1 - b), b2 = (1 - a) * b;
			double *o = samp + 4 * (size_t)k;
#pragma unroll
			for (int c = 0; c < 3; ++c)
				o[c] = b0 * v[c] + b1 * v[3 + c] + b2 * v[6 + c];
			o[3] = b0 * t.e[0] + b1 * t.e[1] + b2 * t.e[2];
		}
		__syncthreads();
	}
	(void)s_item;
}

// one warp per (env, taxel): the environment's samples in order, same per-sample arithmetic as taxel_gather_kernel
__global__ void __launch_bounds__(128) taxel_ai_gather_kernel(TaxelDev td, StepIO io)
{
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const long unit = (long)blockIdx.x * 4 + wib;
	if (unit >= (long)io.n_env * td.n_taxels)
		return;
	const int env = (int)(unit / td.n_taxels), taxel = (int)(unit - (long)env * td.n_taxels);
	float *out  = td.values + unit;
	const int m = td.n_samples[env];
	if (m == 0) { // no sample at all: the message is zeroed (:455-477)
		if (lane == 0)
			*out = 0.0f;
		return;
	}
	const double *R  = io.xmat + ((size_t)env * io.n_geoms + td.geom) * 9;
	const double *xp = io.xpos + ((size_t)env * io.n_geoms + td.geom) * 3;
	double tw[3];
	taxel_world(td, R, xp, taxel, tw);
	const double tsq = tw[0] * tw[0] + tw[1] * tw[1] + tw[2] * tw[2];
	const double margin = td.include_margin, margin_sq = margin * margin, res = td.sample_resolution;
	const double *samp = td.samples + (size_t)env * td.max_samples * 4;
	double pressure = 0, dmin = 1e300, pmin = 0;
	int ws = 0, kmin = 0x7fffffff;
	for (int k = lane; k < m; k += 32) {
		const double *p = samp + 4 * (size_t)k;
		double dj = (-2 * (tw[0] * p[0] + tw[1] * p[1] + tw[2] * p[2]) + tsq) + (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
		double pj = p[3];
		if (td.method == 0) {
			if (dj < dmin)
				dmin = dj, pmin = pj, kmin = k;
		} else if (dj < margin_sq) {
			double w = fmax(0.0, margin - sqrt(dj));
			pressure += (w * w) * fabs(pj);
			ws += 1;
		}
	}
	if (td.method == 0) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			double d2 = __shfl_xor_sync(0xffffffffu, dmin, o), p2 = __shfl_xor_sync(0xffffffffu, pmin, o);
			int k2 = __shfl_xor_sync(0xffffffffu, kmin, o);
			if (d2 < dmin || (d2 == dmin && k2 < kmin))
				dmin = d2, pmin = p2, kmin = k2;
		}
		if (lane == 0 && dmin < margin_sq)
			*out = (td.visualize && fabs(pmin) > 1e-6) ? (float)pmin : 0.0f; // :312-328
		return;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		pressure += __shfl_xor_sync(0xffffffffu, pressure, o);
		ws += __shfl_xor_sync(0xffffffffu, ws, o);
	}
	if (lane == 0 && ws > 0)
		*out = (float)(pressure * res); // :409-412
}

static int launch_taxel_area_importance(const TaxelDev &td, const StepIO &io, const PairDesc *d_pairs, cudaStream_t s)
{
	const int n_tiles = (io.n_env + SCAN_TILE - 1) / SCAN_TILE;
	const int tgrid   = (int)std::max<long>(1, std::min<long>(((long)io.max_tris + 255) / 256, (long)io.n_sms * 8));
	const size_t smem = (size_t)TAXEL_AI_CAP * (sizeof(unsigned long long) + sizeof(double) + sizeof(int));
	// opt in to > 48 KB dynamic shared memory (idempotent and cheap; contexts may live on several devices)
	cudaFuncSetAttribute(taxel_ai_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaMemsetAsync(td.env_tris, 0, (size_t)io.n_env * sizeof(int32_t), s);
	cudaMemsetAsync(td.env_cursor, 0, (size_t)io.n_env * sizeof(int32_t), s);
	taxel_ai_list_kernel<false><<<tgrid, 256, 0, s>>>(td, io, d_pairs);
	scan_tiles_kernel<<<n_tiles, 256, 0, s>>>(td.env_tris, td.env_offset, td.scan_tmp, io.n_env);
	scan_sums_kernel<<<1, 1024, 0, s>>>(td.scan_tmp, n_tiles, td.env_offset + io.n_env);
	scan_add_kernel<<<n_tiles, 256, 0, s>>>(td.env_offset, td.scan_tmp, io.n_env);
	taxel_ai_list_kernel<true><<<tgrid, 256, 0, s>>>(td, io, d_pairs);
	taxel_ai_sample_kernel<<<io.n_env, 256, smem, s>>>(td, io);
	taxel_ai_gather_kernel<<<(io.n_env * td.n_taxels + 3) / 4, 128, 0, s>>>(td, io);
	return 7;
}

int launch_taxel(const TaxelDev &td, const StepIO &io, const PairDesc *d_pairs, cudaStream_t s)
{
	if (td.sample_method == 1) {
		if (io.n_env * td.n_taxels <= 0 || io.max_tris <= 0 || !io.tri_vd || !io.tri_elem)
			return 0;
		return launch_taxel_area_importance(td, io, d_pairs, s);
	}
	const int ncell = io.n_env * td.n_taxels;
	if (ncell <= 0)
		return 0;
	const int n_tiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
	const int tgrid   = (int)std::max<long>(1, std::min<long>(((long)io.max_tris + 255) / 256, (long)io.n_sms * 8));
	taxel_clear_kernel<<<(std::max(ncell, io.n_env) + 255) / 256, 256, 0, s>>>(td, ncell, io.n_env);
	if (io.max_tris > 0 && io.tri_vd)
		taxel_bin_kernel<false><<<tgrid, 256, 0, s>>>(td, io, d_pairs);
	scan_tiles_kernel<<<n_tiles, 256, 0, s>>>(td.bin_count, td.bin_offset, td.scan_tmp, ncell);
	scan_sums_kernel<<<1, 1024, 0, s>>>(td.scan_tmp, n_tiles, td.bin_offset + ncell);
	scan_add_kernel<<<n_tiles, 256, 0, s>>>(td.bin_offset, td.scan_tmp, ncell);
	if (io.max_tris > 0 && io.tri_vd)
		taxel_bin_kernel<true><<<tgrid, 256, 0, s>>>(td, io, d_pairs);
	taxel_gather_kernel<<<(ncell + 3) / 4, 128, 0, s>>>(td, io);
	return 7;
}

} // namespace hcs
