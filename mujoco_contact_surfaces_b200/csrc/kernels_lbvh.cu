// K2 lbvh_build — GPU LBVH over the tets of a soft geom (north star stage 2: "GPU LBVH built on Morton
// codes").  Replaces the Bvh<Obb, VolumeMesh> constructor the reference calls at
// mujoco_contact_surfaces_plugin.cpp:656 (and :680, :704, :730, :790).  Built once per geom in the geom's own
// frame at hcs_finalize / hcs_update_geom: meshes are rigid, so nothing is rebuilt inside the step.
//
//   lbvh_leaf_kernel     1 thread / tet: leaf box (fp64), centroid, 30-bit Morton code -> 64-bit key (code, tet)
//   bitonic_step_kernel  in-place bitonic sort of the keys (load time only; n <= a few 100 k)
//   lbvh_karras_kernel   1 thread / internal node: Karras 2012 radix tree on the sorted unique keys
//   lbvh_refit_kernel    1 thread / leaf: bottom-up boxes with one atomic counter per internal node; the second
//                        arrival owns the node, writes its 64-B record (both children's boxes as floats rounded
//                        outward + child ids) and climbs on
// The result is bit-identical to the host builder kept in engine.cu for cross-checking (HCS_LBVH_HOST=1).
#include "hcs_internal.h"

namespace hcs {

__device__ __forceinline__ unsigned expand_bits10(unsigned v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

__global__ void __launch_bounds__(128) lbvh_leaf_kernel(GeomDev g, double3 glo, double3 ghi, double *leaf_box,
                                                        unsigned long long *keys, int n_pad)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_pad)
		return;
	if (t >= g.n_elems) {
		keys[t] = ~0ull; // padding sorts to the end
		return;
	}
	int4 idx  = reinterpret_cast<const int4 *>(g.elems)[t];
	int vi[4] = { idx.x, idx.y, idx.z, idx.w };
	double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 }, c[3] = { 0, 0, 0 };
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const double *p = g.verts + 3 * (size_t)vi[k];
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			lo[a] = fmin(lo[a], p[a]);
			hi[a] = fmax(hi[a], p[a]);
			c[a] += 0.25 * p[a];
		}
	}
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		leaf_box[6 * (size_t)t + a]     = lo[a];
		leaf_box[6 * (size_t)t + 3 + a] = hi[a];
	}
	const double gl[3] = { glo.x, glo.y, glo.z }, gh[3] = { ghi.x, ghi.y, ghi.z };
	unsigned q[3];
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		double ext = gh[a] - gl[a];
		double u   = ext > 0 ? (c[a] - gl[a]) / ext : 0.0;
		q[a]       = (unsigned)fmin(1023.0, fmax(0.0, u * 1024.0));
	}
	unsigned code = expand_bits10(q[0]) * 4 + expand_bits10(q[1]) * 2 + expand_bits10(q[2]);
	keys[t]       = ((unsigned long long)code << 32) | (unsigned)t;
}

__global__ void __launch_bounds__(256) bitonic_step_kernel(unsigned long long *a, int j, int k, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	int ixj = i ^ j;
	if (ixj > i) {
		unsigned long long x = a[i], y = a[ixj];
		bool ascending = (i & k) == 0;
		if ((x > y) == ascending) {
			a[i]   = y;
			a[ixj] = x;
		}
	}
}

__device__ __forceinline__ int key_delta(const unsigned long long *keys, int n, int i, int j)
{
	if (j < 0 || j >= n)
		return -1;
	return __clzll((long long)(keys[i] ^ keys[j]));
}

// children: >= 0 internal node, < 0 ~(sorted leaf slot); parent[] indexed by internal node / (n-1 + leaf slot)
__global__ void __launch_bounds__(128) lbvh_karras_kernel(const unsigned long long *keys, int n, int *left, int *right,
                                                          int *parent)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1)
		return;
	int d    = key_delta(keys, n, i, i + 1) - key_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
	int dmin = key_delta(keys, n, i, i - d);
	int lmax = 2;
	while (key_delta(keys, n, i, i + lmax * d) > dmin)
		lmax *= 2;
	int l = 0;
	for (int t = lmax / 2; t >= 1; t /= 2)
		if (key_delta(keys, n, i, i + (l + t) * d) > dmin)
			l += t;
	int j     = i + l * d;
	int dnode = key_delta(keys, n, i, j);
	int s     = 0;
	for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
		if (key_delta(keys, n, i, i + (s + t) * d) > dnode)
			s += t;
		if (t == 1)
			break;
	}
	int gamma = i + s * d + min(d, 0);
	int lc = min(i, j) == gamma ? ~gamma : gamma;
	int rc = max(i, j) == gamma + 1 ? ~(gamma + 1) : gamma + 1;
	left[i]  = lc;
	right[i] = rc;
	parent[lc < 0 ? (n - 1) + ~lc : lc] = i;
	parent[rc < 0 ? (n - 1) + ~rc : rc] = i;
	if (i == 0)
		parent[0] = -1;
}

__global__ void __launch_bounds__(128) lbvh_refit_kernel(const unsigned long long *keys, int n, const int *left,
                                                         const int *right, const int *parent, const double *leaf_box,
                                                         double *node_box, int *arrived, BvhNode *nodes)
{
	int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n)
		return;
	int node = parent[(n - 1) + slot];
	while (node >= 0) {
		__threadfence();
		if (atomicAdd(arrived + node, 1) == 0)
			return; // the sibling subtree is not finished yet: its thread will own this node
		int child[2] = { left[node], right[node] };
		double box[2][6];
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			const double *src = child[c] < 0 ? leaf_box + 6 * (size_t)(unsigned)(keys[~child[c]] & 0xffffffffu) :
			                                    node_box + 6 * (size_t)child[c];
#pragma unroll
			for (int a = 0; a < 6; ++a)
				box[c][a] = __ldcg(src + a); // written by another thread: read through L2
		}
		BvhNode nd;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			nd.llo[a] = __double2float_rd(box[0][a]);
			nd.lhi[a] = __double2float_ru(box[0][3 + a]);
			nd.rlo[a] = __double2float_rd(box[1][a]);
			nd.rhi[a] = __double2float_ru(box[1][3 + a]);
			node_box[6 * (size_t)node + a]     = fmin(box[0][a], box[1][a]);
			node_box[6 * (size_t)node + 3 + a] = fmax(box[0][3 + a], box[1][3 + a]);
		}
		nd.left   = child[0] < 0 ? ~(int)(unsigned)(keys[~child[0]] & 0xffffffffu) : child[0];
		nd.right  = child[1] < 0 ? ~(int)(unsigned)(keys[~child[1]] & 0xffffffffu) : child[1];
		nd.pad[0] = nd.pad[1] = 0;
		nodes[node] = nd;
		node = parent[node];
	}
}

// scratch: leaf_box 6n doubles, node_box 6(n-1) doubles, keys n_pad u64, left/right (n-1) ints, parent (2n-1) ints,
// arrived (n-1) ints — allocated by the caller (engine.cu), sizes from lbvh_scratch_bytes().
size_t lbvh_scratch_bytes(int n)
{
	int n_pad = 1;
	while (n_pad < n)
		n_pad <<= 1;
	size_t b = 0;
	b += sizeof(double) * 6 * (size_t)n;           // leaf_box
	b += sizeof(double) * 6 * (size_t)n;           // node_box
	b += sizeof(unsigned long long) * (size_t)n_pad; // keys
	b += sizeof(int) * (size_t)n * 5;              // left, right, parent (2n), arrived
	return b + 256;
}

void launch_bitonic_sort_u64(unsigned long long *keys, int n_pad, cudaStream_t s)
{
	for (int k = 2; k <= n_pad; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1)
			bitonic_step_kernel<<<(n_pad + 255) / 256, 256, 0, s>>>(keys, j, k, n_pad);
}

void launch_build_lbvh(const GeomDev &g, const double glo[3], const double ghi[3], void *scratch, cudaStream_t s)
{
	int n = g.n_elems;
	if (n < 2)
		return; // single-tet trees are written by the host
	int n_pad = 1;
	while (n_pad < n)
		n_pad <<= 1;
	char *p = static_cast<char *>(scratch);
	double *leaf_box = reinterpret_cast<double *>(p);
	p += sizeof(double) * 6 * (size_t)n;
	double *node_box = reinterpret_cast<double *>(p);
	p += sizeof(double) * 6 * (size_t)n;
	unsigned long long *keys = reinterpret_cast<unsigned long long *>(p);
	p += sizeof(unsigned long long) * (size_t)n_pad;
	int *left = reinterpret_cast<int *>(p), *right = left + n, *parent = right + n, *arrived = parent + 2 * (size_t)n;
	cudaMemsetAsync(arrived, 0, sizeof(int) * (size_t)n, s);
	lbvh_leaf_kernel<<<(n_pad + 127) / 128, 128, 0, s>>>(g, make_double3(glo[0], glo[1], glo[2]),
	                                                      make_double3(ghi[0], ghi[1], ghi[2]), leaf_box, keys, n_pad);
	launch_bitonic_sort_u64(keys, n_pad, s);
	lbvh_karras_kernel<<<(n - 1 + 127) / 128, 128, 0, s>>>(keys, n, left, right, parent);
	lbvh_refit_kernel<<<(n + 127) / 128, 128, 0, s>>>(keys, n, left, right, parent, leaf_box, node_box, arrived, g.nodes);
}

} // namespace hcs
