// K0 mesh generation on the GPU ("GPU mesh/field/LBVH (re)build").
//
// Replaces the host side of MakeSphereVolumeMesh / MakeEllipsoidVolumeMesh + MakeSphereSurfaceMesh as the plugin calls them
// (mujoco_contact_surfaces_plugin.cpp:650-700): the refined-octahedron unit sphere (DESIGN.md "Mesh specification": vertex
// 0 = centre, every refinement splits each boundary triangle (a,b,c) into (a,ab,ca),(ab,b,bc),(ca,bc,c),(ab,bc,ca), midpoints
// numbered in order of first use and re-projected onto the sphere).  "First use" is what makes the numbering sequential on
// the host; here one refinement level is
//   edge_keys      one thread per (triangle, edge) slot: key = (min vertex, max vertex, slot)
//   bitonic sort   the two slots of an edge become neighbours, the first-use slot in front          (kernels_lbvh.cu)
//   edge_owner     the first-use slot owns the midpoint; flags in slot order
//   exclusive scan rank of the owner among the owners = position of its midpoint in first-use order   (kernels_tactile.cu)
//   edge_midpoint  owners create their vertex ((A + B) / 2 normalised, the host's expression tree, IEEE sqrt and division)
//   subdivide      one thread per triangle writes its four children
// so vertex ids, element order and every coordinate are bit-identical to mesh_host.cpp (engine.cu checks it when
// HCS_MESHGEN_CHECK is set; tests/test_gpu_parity.py::test_gpu_sphere_generation_equals_host_generator).
#include "dmath.cuh"
#include "hcs_internal.h"

namespace hcs {

// 21 + 21 + 22 bits: up to 2 M vertices and 4 M edge slots (refinement level 8: 393 218 vertices, 1 572 864 slots)
constexpr int MG_SLOT_BITS = 22, MG_VERT_BITS = 21;
constexpr unsigned long long MG_SLOT_MASK = (1ull << MG_SLOT_BITS) - 1ull;

__global__ void __launch_bounds__(256) mg_edge_keys_kernel(const int32_t *tri, int n_slots, unsigned long long *keys, int n_pad)
{
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n_pad)
		return;
	if (slot >= n_slots) {
		keys[slot] = ~0ull; // padding sorts to the end
		return;
	}
	const int t = slot / 3, k = slot - 3 * t;
	const int p = tri[3 * t + k], q = tri[3 * t + (k == 2 ? 0 : k + 1)];
	const unsigned long long lo = (unsigned long long)min(p, q), hi = (unsigned long long)max(p, q);
	keys[slot] = (lo << (MG_VERT_BITS + MG_SLOT_BITS)) | (hi << MG_SLOT_BITS) | (unsigned long long)slot;
}

// sorted keys: positions 2j and 2j + 1 hold the two slots of edge j, the earlier slot first
__global__ void __launch_bounds__(256) mg_edge_owner_kernel(const unsigned long long *keys, int n_edges, int32_t *owner_flag,
                                                            int32_t *partner, int32_t *bad)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_edges)
		return;
	const unsigned long long k0 = keys[2 * j], k1 = keys[2 * j + 1];
	if ((k0 >> MG_SLOT_BITS) != (k1 >> MG_SLOT_BITS))
		atomicOr(bad, 1); // not a closed two-manifold surface (cannot happen for the refined octahedron)
	const int s0 = (int)(k0 & MG_SLOT_MASK), s1 = (int)(k1 & MG_SLOT_MASK);
	owner_flag[s0] = 1, partner[s0] = s0;
	owner_flag[s1] = 0, partner[s1] = s0;
}

__global__ void __launch_bounds__(256) mg_edge_midpoint_kernel(const int32_t *tri, int n_slots, const int32_t *owner_flag,
                                                               const int32_t *partner, const int32_t *rank, int n_verts, double *verts,
                                                               int32_t *edge_id)
{
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n_slots)
		return;
	const int id  = n_verts + rank[partner[slot]];
	edge_id[slot] = id;
	if (!owner_flag[slot])
		return;
	const int t = slot / 3, k = slot - 3 * t;
	const int p = tri[3 * t + k], q = tri[3 * t + (k == 2 ? 0 : k + 1)];
	const double *A = verts + 3 * (size_t)min(p, q), *B = verts + 3 * (size_t)max(p, q);
	D3 m           = mk((A[0] + B[0]) * 0.5, (A[1] + B[1]) * 0.5, (A[2] + B[2]) * 0.5);
	const double z = dot(m, m);
	if (z > 0) {
		const double s = sqrt(z);
		m              = mk(m.x / s, m.y / s, m.z / s);
	}
	verts[3 * (size_t)id] = m.x, verts[3 * (size_t)id + 1] = m.y, verts[3 * (size_t)id + 2] = m.z;
}

__global__ void __launch_bounds__(256) mg_subdivide_kernel(const int32_t *tri, int n_tri, const int32_t *edge_id, int32_t *next)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_tri)
		return;
	const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
	const int ab = edge_id[3 * t], bc = edge_id[3 * t + 1], ca = edge_id[3 * t + 2];
	const int four[12] = { a, ab, ca, ab, b, bc, ca, bc, c, ab, bc, ca };
#pragma unroll
	for (int k = 0; k < 12; ++k)
		next[12 * (size_t)t + k] = four[k];
}

// soft: tets (centre, a, b, c) in boundary-triangle order; rigid: the boundary triangles with the centre vertex dropped
__global__ void __launch_bounds__(256) mg_sphere_elems_kernel(const int32_t *tri, int n_tri, int soft, int32_t *elems)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_tri)
		return;
	const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
	if (soft) {
		elems[4 * (size_t)t] = 0, elems[4 * (size_t)t + 1] = a, elems[4 * (size_t)t + 2] = b, elems[4 * (size_t)t + 3] = c;
	} else {
		elems[3 * (size_t)t] = a - 1, elems[3 * (size_t)t + 1] = b - 1, elems[3 * (size_t)t + 2] = c - 1;
	}
}

int unit_sphere_max_gpu_level() { return 8; }
void unit_sphere_counts(int level, int *n_verts, int *n_tri)
{
	long nt = 8, nv = 7;
	for (int l = 0; l < level; ++l) {
		nv += 3 * nt / 2;
		nt *= 4;
	}
	*n_verts = (int)nv, *n_tri = (int)nt;
}

// verts: [n_verts][3] doubles, tri: [n_tri][3] ints (device, sized by unit_sphere_counts); scratch: caller-allocated, at least
// unit_sphere_scratch_bytes(level).  Returns false when the surface was not closed (flag read back by the caller).
size_t unit_sphere_scratch_bytes(int level)
{
	int nv, nt;
	unit_sphere_counts(level, &nv, &nt);
	size_t slots = 3 * (size_t)nt / 4 + 16; // the last level refines nt / 4 triangles
	size_t pad   = 1;
	while (pad < slots)
		pad <<= 1;
	return pad * 8 + slots * 4 * 4 + (slots / 1024 + 4) * 4 + (size_t)nt * 3 * 4 + 1024;
}

void launch_unit_sphere(int level, double *verts, int32_t *tri, void *scratch, int32_t *bad_flag, cudaStream_t s)
{
	static const double v0[7][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 }, { -1, 0, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } };
	static const int32_t t0[24]  = { 1, 2, 5, 2, 3, 5, 3, 4, 5, 4, 1, 5, 2, 1, 6, 3, 2, 6, 4, 3, 6, 1, 4, 6 };
	int nv_final, nt_final;
	unit_sphere_counts(level, &nv_final, &nt_final);
	// two triangle buffers: the result must end in `tri`, so an even number of levels starts there
	char *p          = static_cast<char *>(scratch);
	int32_t *tri_alt = reinterpret_cast<int32_t *>(p);
	p += (size_t)nt_final * 3 * 4;
	p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(p) + 255) & ~(uintptr_t)255);
	size_t max_slots = 3 * (size_t)nt_final / 4 + 16, pad_max = 1;
	while (pad_max < max_slots)
		pad_max <<= 1;
	unsigned long long *keys = reinterpret_cast<unsigned long long *>(p);
	p += pad_max * 8;
	int32_t *owner = reinterpret_cast<int32_t *>(p), *partner = owner + max_slots, *rank = partner + max_slots, *edge_id = rank + max_slots;
	int32_t *tile_tmp = edge_id + max_slots;
	int32_t *cur = (level % 2 == 0) ? tri : tri_alt, *nxt = (level % 2 == 0) ? tri_alt : tri;
	cudaMemcpyAsync(verts, v0, sizeof v0, cudaMemcpyHostToDevice, s);
	cudaMemcpyAsync(cur, t0, sizeof t0, cudaMemcpyHostToDevice, s);
	cudaMemsetAsync(bad_flag, 0, sizeof(int32_t), s);
	int nv = 7, nt = 8;
	for (int l = 0; l < level; ++l) {
		const int n_slots = 3 * nt;
		int n_pad         = 1;
		while (n_pad < n_slots)
			n_pad <<= 1;
		mg_edge_keys_kernel<<<(n_pad + 255) / 256, 256, 0, s>>>(cur, n_slots, keys, n_pad);
		launch_bitonic_sort_u64(keys, n_pad, s);
		mg_edge_owner_kernel<<<(n_slots / 2 + 255) / 256, 256, 0, s>>>(keys, n_slots / 2, owner, partner, bad_flag);
		launch_exclusive_scan_i32(owner, rank, tile_tmp, n_slots, tile_tmp + n_slots / 1024 + 2, s);
		mg_edge_midpoint_kernel<<<(n_slots + 255) / 256, 256, 0, s>>>(cur, n_slots, owner, partner, rank, nv, verts, edge_id);
		mg_subdivide_kernel<<<(nt + 255) / 256, 256, 0, s>>>(cur, nt, edge_id, nxt);
		nv += n_slots / 2;
		nt *= 4;
		std::swap(cur, nxt);
	}
}

void launch_sphere_elems(const int32_t *tri, int n_tri, int soft, int32_t *elems, cudaStream_t s)
{
	mg_sphere_elems_kernel<<<(n_tri + 255) / 256, 256, 0, s>>>(tri, n_tri, soft, elems);
}

} // namespace hcs
