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
template <bool FILL>
__global__ void __launch_bounds__(256) tactile_bin_kernel(SensorDev sd, StepIO io, const PairDesc *pairs)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int n = min(*io.tri_count, io.max_tris);
	if (i >= n)
		return;
	const TactileTri &t = io.tri_pool[i];
	const PairDesc &P   = pairs[t.pair];
	if (P.gM != sd.geom && P.gN != sd.geom)
		return;
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

constexpr int RASTER_WARPS = 4;
constexpr int MAX_SAMPLES  = 32 * 32;

__global__ void __launch_bounds__(32 * RASTER_WARPS) tactile_raster_kernel(SensorDev sd, StepIO io)
{
	__shared__ float tile[RASTER_WARPS][MAX_SAMPLES];
	int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int ntax = sd.cx * sd.cy;
	long unit = (long)blockIdx.x * RASTER_WARPS + wib;
	if (unit >= (long)io.n_env * ntax)
		return;
	int env = (int)(unit / ntax), cell = (int)(unit - (long)env * ntax);
	int x = cell % sd.cx, y = cell / sd.cx;
	// the reference stores taxel (x,y) at x + cy*y (flat_tactile_sensor.cpp:396-397; quirk Q8: only a
	// bijection when cx == cy); out-of-range indices of non-square arrays are dropped
	int oidx = x + sd.cy * y;
	if (oidx >= ntax)
		return;
	int first = sd.bin_offset[unit];
	int n     = min(sd.bin_count[unit], max(sd.items_cap - first, 0));
	float *out = sd.image + (size_t)env * ntax + oidx;
	if (n == 0) {
		if (lane == 0)
			*out = 0.0f;
		return;
	}
	const int32_t *items = sd.bin_items + first;
	const double *R  = io.xmat + ((size_t)env * io.n_geoms + sd.geom) * 9;
	const double *xp = io.xpos + ((size_t)env * io.n_geoms + sd.geom) * 3;
	double rot[9];
#pragma unroll
	for (int k = 0; k < 9; ++k)
		rot[k] = R[k];
	float xs = (float)sd.size[0], ys = (float)sd.size[1], zs = (float)sd.size[2];
	F3 normal      = f3((float)rot[2], (float)rot[5], (float)rot[8]);
	double topleft[3] = { (double)(-xs), (double)(-ys), (double)zs };
	double resolution = sd.resolution;
	float rS = sd.rS, rmean = sd.rmean;
	int S  = sd.S;
	F3 D   = f3(-normal.x, -normal.y, -normal.z);
	F3 off = normal * (float)1e-8;
	double tmax = 1.5 * zs;
	for (int s = lane; s < S * S; s += 32) {
		int i = s / S, j = s - i * S;
		double pos[3] = { topleft[0] + x * resolution + (double)((float)i * rS) + 0.5 * (double)rS,
			              topleft[1] + y * resolution + (double)((float)j * rS) + 0.5 * (double)rS, 1.5 * (double)zs };
		double w[3]   = { rot[0] * pos[0] + rot[1] * pos[1] + rot[2] * pos[2], rot[3] * pos[0] + rot[4] * pos[1] + rot[5] * pos[2],
			              rot[6] * pos[0] + rot[7] * pos[1] + rot[8] * pos[2] };
		w[0] += xp[0], w[1] += xp[1], w[2] += xp[2];
		F3 O = f3((float)w[0], (float)w[1], (float)w[2]) + off;
		float best_t = 1e30f, best_u = 0, best_v = 0;
		int best = -1, best_pair = 0, best_order = 0;
		for (int k = 0; k < n; ++k) {
			int ti = items[k];
			const TactileTri &t = io.tri_pool[ti];
			F3 v0 = f3(t.v[0], t.v[1], t.v[2]), v1 = f3(t.v[3], t.v[4], t.v[5]), v2 = f3(t.v[6], t.v[7], t.v[8]);
			// Moeller-Trumbore, bvh.cpp:49-74
			F3 edge1 = v1 - v0, edge2 = v2 - v0;
			F3 h    = crossf(D, edge2);
			float a = dotf(edge1, h);
			if (fabsf(a) < 1e-10f)
				continue;
			float f = 1.0f / a;
			F3 sv   = O - v0;
			float u = f * dotf(sv, h);
			if (u < 0.0f || u > 1.0f)
				continue;
			F3 q    = crossf(sv, edge1);
			float v = f * dotf(D, q);
			if (v < 0.0f || u + v > 1.0f)
				continue;
			float tt = f * dotf(edge2, q);
			if (!(tt > 0.0f))
				continue;
			bool better = tt < best_t;
			if (tt == best_t && best >= 0) // deterministic tie-break independent of pool order
				better = t.pair < best_pair || (t.pair == best_pair && t.order < best_order);
			if (better) {
				best_t = tt, best_u = u, best_v = v, best = ti;
				best_pair = t.pair, best_order = t.order;
			}
		}
		float val = 0.0f;
		if (best >= 0 && (double)best_t < tmax && best_t > 0.0f) {
			const TactileTri &t = io.tri_pool[best];
			double b0 = (double)(1 - best_u - best_v), b1 = (double)best_u, b2 = (double)best_v;
			double ev = b0 * t.e[0];
			ev += b1 * t.e[1];
			ev += b2 * t.e[2];
			float raw = (float)(ev * (double)rmean);
			val       = sd.weights[s] * raw;
		}
		tile[wib][s] = val;
	}
	__syncwarp();
	if (lane == 0) {
		float avg = 0;
		for (int s = 0; s < S * S; ++s)
			avg += tile[wib][s];
		*out = avg;
	}
}

__global__ void tactile_clear_kernel(SensorDev sd, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		sd.bin_count[i]  = 0;
		sd.bin_cursor[i] = 0;
	}
}

// 7 launches: clear, count, 3-phase scan, fill, raster
int launch_tactile(const SensorDev &sd, const StepIO &io, const PairDesc *d_pairs, cudaStream_t s)
{
	int ncell   = io.n_env * sd.cx * sd.cy;
	int n_tiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
	int tgrid   = (io.max_tris + 255) / 256;
	tactile_clear_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(sd, ncell);
	if (io.max_tris > 0)
		tactile_bin_kernel<false><<<tgrid, 256, 0, s>>>(sd, io, d_pairs);
	scan_tiles_kernel<<<n_tiles, 256, 0, s>>>(sd.bin_count, sd.bin_offset, sd.scan_tmp, ncell);
	scan_sums_kernel<<<1, 1024, 0, s>>>(sd.scan_tmp, n_tiles, sd.bin_offset + ncell);
	scan_add_kernel<<<n_tiles, 256, 0, s>>>(sd.bin_offset, sd.scan_tmp, ncell);
	if (io.max_tris > 0)
		tactile_bin_kernel<true><<<tgrid, 256, 0, s>>>(sd, io, d_pairs);
	tactile_raster_kernel<<<(ncell + RASTER_WARPS - 1) / RASTER_WARPS, 32 * RASTER_WARPS, 0, s>>>(sd, io);
	return 5 + (io.max_tris > 0 ? 2 : 0);
}

} // namespace hcs
