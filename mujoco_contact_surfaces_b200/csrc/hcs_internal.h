// Internal device data layout + kernel launchers of libhcs_b200 (sm_100a).
// All device arithmetic is compiled with -fmad=false: the classification decisions of the
// clipping kernels (signed distance <= 0, duplicate threshold 1e-14) must reproduce the fp64
// operation order of the restated Drake path bit for bit (DESIGN.md "Floating-point contract").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef __CUDACC__
#include <map>
#include <mutex>
#include <utility>
#endif

#include "../../include/hcs.h"

namespace hcs {

// ---- resident geometry records (HBM) ---------------------------------------------------------
// Gather records are made of whole 32-byte groups and 32-byte aligned so that one thread fetches its element with
// 256-bit loads (dmath.cuh ld4); consecutive lanes of the streaming kernels read consecutive records.

struct __align__(32) TetGeom { // 128 B = one cache line = 4 x 32 B: tet vertices (3 groups) + vertex pressures (1)
	double v[4][3];
	double e[4];
};
struct __align__(32) TetField { // 192 B = 6 x 32 B: what the tet contributes to a (tet, triangle) clip
	double plane[4][4]; // outward unit normal + offset of faces {1,2,3},{0,3,2},{0,1,3},{0,2,1}
	double grad[3];     // pressure gradient
	double e0;          // pressure at the geom-frame origin
	double ghat[3];     // normalized gradient (cull direction)
	double pad;
};
struct __align__(32) TetLeaf32 { // 128 B = one line: float copy of what the broadphase leaf filter reads (records.cuh)
	float plane[4][4]; // as TetField::plane, rounded to nearest
	float ghat[3];
	float pad;
	float v[4][3]; // tet vertices
};
struct __align__(32) TetLeafSS32 { // 96 B: float copy of what the soft-soft leaf filter reads of the tree-side tet
	float grad[3], e0; // as TetField::grad / e0
	float ghat[3], pad;
	float v[4][3];
	float pad2[4];
};
struct __align__(32) TetBox32 { // 32 B: the tet's box in the geom frame, rounded outward (broadphase sweep of small geoms)
	float lo[3], hi[3];
	float pad[2];
};
struct __align__(32) TriRec { // 96 B = 3 x 32 B: rigid triangle vertices + unit normal
	double v[3][3];
	double n[3];
};
struct __align__(32) BvhNode { // 64 B = 2 x 32 B: both children's boxes live in the parent
	float llo[3], lhi[3], rlo[3], rhi[3];
	int32_t left, right; // >= 0 internal node, < 0 leaf holding element ~child
	float pad[2];
};

struct GeomDev {
	int kind; // 0 rigid mesh, 1 soft, 2 plane
	int n_verts, n_elems;
	double *verts;      // [nv][3]
	int32_t *elems;     // [ne][4|3]
	double *pressure;   // [nv]
	TetGeom *tet_geom;  // soft
	TetField *tet_field;
	TetLeaf32 *tet_leaf32;
	TetLeafSS32 *tet_leafss32;
	TetBox32 *tet_box32;
	TriRec *tris;       // rigid
	BvhNode *nodes;     // soft: LBVH over tets (root = node 0)
	double bound_c[3];  // bounding sphere in the geom frame
	double bound_r;
	float root_lo[3], root_hi[3]; // soft: box of the whole LBVH
	// Per-environment sizes (hcs_set_env_sizes: domain-randomised geometry): every environment has
	// its own vertices, pressures, element records and LBVH of ONE shared topology, laid out env-major.  env_stride = element
	// records per environment (0: one geometry for all environments, the arrays above are it), env_stride_nodes likewise
	// for the LBVH, env_stride_verts for verts / pressure; env_bounds = [n_env][ENV_BOUNDS] doubles: bound_c, bound_r,
	// root_lo, root_hi of each environment (the members above then describe environment 0).
	int env_stride, env_stride_nodes, env_stride_verts;
	const double *env_bounds;
	// Large trees: the nodes at depth SPLIT_DEPTH (and the leaves above it), so that the flat traversal can hand the K
	// subtrees of one query to different warps (kernels_broadphase.cu); NULL for small trees and per-environment geometry
	const int32_t *split_nodes; // >= 0 internal node, < 0 leaf ~tet
	int n_split;
#ifdef __CUDACC__
	__host__ __device__ __forceinline__ size_t eoff(int env) const { return (size_t)env * (size_t)env_stride; }
	__host__ __device__ __forceinline__ size_t noff(int env) const { return (size_t)env * (size_t)env_stride_nodes; }
#endif
};
constexpr int ENV_BOUNDS = 10;
constexpr int SPLIT_DEPTH = 6, SPLIT_MIN_TREE = 4096; // subtree split of the flat traversal: 64 subtrees for trees of >= 4096 tets

enum PairKind { PAIR_NONE = 0, PAIR_SOFT_RIGID = 1, PAIR_SOFT_PLANE = 2, PAIR_SOFT_SOFT = 3 };

struct TactileTri { // kTriangle contact-surface triangle handed to the tactile stage (72 B)
	float v[9];
	int32_t env;
	double e[3];
	uint32_t key_hi; // pair << TRI_PAIR_SHIFT | query element >> 2
	uint32_t key_lo; // (query element & 3) << 30 | tree element << 3 | fan triangle
	// (key_hi, key_lo) = (pair, query element of B, tree element of A, fan triangle) is a canonical total order of the
	// triangles of one environment: it does not depend on the order in which candidates were found or clipped
};
constexpr int TRI_PAIR_SHIFT = 22;   // <= 2^10 pairs, <= 2^24 query elements, <= 2^27 tree elements
constexpr int TRI_SLICE_BITS = TRI_PAIR_SHIFT; // (name kept for the pair-count limit in hcs_set_pairs)

// ---- exact per-(env, pair) accumulators ------------------------------------------------------------------
// Every contribution of a contact polygon (F, tau, area, area * centroid) is added to its (env, pair) record as a
// two-limb fixed-point integer (hi: units of 2^-36, lo: units of 2^-80 of the SI value) with 64-bit integer atomics.
// Integer addition is associative and commutative, so the sums do not depend on the order in which candidates are
// found, clipped or added: results are bit-reproducible whatever the batch composition, grid size or scheduling, without
// any ordered per-candidate records.  Range: |scaled sum| < 2^26, resolution 2^-80, up to 2^19 contributions per
// record before the low limb could overflow.
// Per pair the values are first scaled by a power of two (PairDesc::acc_scale, exact) chosen from the pair's size and
// stiffness, so that the range follows the scene: metres and newtons, or millimetre fingertips, or a 1000x model.
constexpr int ACC_LIMBS = 20;   // 10 components x (hi, lo)
constexpr int ACC_COUNTS = 20;  // faces (bits 0..21) | polygons (22..42) | force points (43..63) of the pair's surface
constexpr int ACC_NCLIPPED = 21; // candidates that reached the clipper
constexpr int ACC_NEVALS = 22;  // pair-evals started in the broadphase: LBVH leaf hits / tets classified
constexpr int ACC_WORDS = 24;   // int64 words per (env, pair) record (192 B)
constexpr int ACC_POLY_SHIFT = 22, ACC_POINT_SHIFT = 43;
constexpr double ACC_HI_SCALE = 0x1p36, ACC_LO_SCALE = 0x1p80;

// Everything a step kernel needs to know about one configured geom pair (passed by value).
struct PairDesc {
	int kind;
	int index;           // pair index inside the scene
	int gA, gB;          // gA: soft geom whose frame hosts the computation; gB: the other one
	int gM, gN;          // min / max configuration index
	double sign;         // +1 when gA == gM, -1 otherwise (ContactSurface M/N swap)
	double dissipation;  // calcCombinedDissipation (plugin.cpp:138-159)
	double mu;           // combined dynamic friction
	int nq, n_tree;      // query elements (of gB) / tree elements (tets of gA); plane: nq = tets of gA
	int n_slices, slice_q;
	int emit_tactile;    // pair touches a sensor geom and representation is kTriangle
	GeomDev A, B;
	uint8_t *nverts;       // [contrib_cap] polygon vertex count of flat-list candidate g (diagnostics: hcs_get_emitted)
	// ---- candidates (all three pair kinds): ONE flat list per pair for the whole batch, in no particular order.  A
	// broadphase warp stages the candidates of its units in shared memory and appends them with one atomicAdd.
	uint4 *flat;           // [contrib_cap] (query element, tree element | skip mask << 28, env, 0)
	int contrib_cap;
	int64_t *accum;        // [n_env][ACC_WORDS] exact accumulators (above), zero between steps (the finalize kernel
	                       // clears what it has read)
	double acc_scale, acc_unscale; // 2^-k / 2^k: contributions are scaled into the accumulators' range (exact)
	int32_t *counters;     // [0] candidates in the flat list, [1] next 32-candidate chunk of the narrowphase,
	                       // [2] next (env, slice) unit / next batch of alive queries of the broadphase, [3] alive
	                       // queries (flat broadphase); zeroed per step
	float *alive;          // [alive_cap][ALIVE_WORDS] query elements that passed the root test (kernels_broadphase.cu)
	int alive_cap;
	double *pair_ctx;      // [n_env][PAIR_CTX_DOUBLES] poses, velocities, X_AB written by the broadphase
};
constexpr int PAIR_COUNTERS = 4;
constexpr int ALIVE_WORDS   = 32; // floats per alive-query record (128 B)
// candidate record .y = tree element | skip mask << 28: planes of the tet the clip may leave out (narrow.cuh cand_tet_tri)
constexpr int CAND_MASK_SHIFT      = 28;
constexpr unsigned CAND_ELEM_MASK  = (1u << CAND_MASK_SHIFT) - 1u;

constexpr int PAIR_CTX_DOUBLES = 48;
// layout of one context block, in 32-byte groups: R_WA[9] xA[3] | R_AB[9] p_AB[3] | wA[3] - | vA[3] - | xB[3] - | wB[3] - |
// vB[3] - | p_BAo[3] -
constexpr int CTX_XA = 9, CTX_RAB = 12, CTX_PAB = 21, CTX_WA = 24, CTX_VA = 28, CTX_XB = 32, CTX_WB = 36, CTX_VB = 40,
              CTX_PBA = 44;

struct StepIO {
	int n_env, n_geoms, n_pairs;
	int n_sms; // SMs of the device: persistent grids are sized in multiples of it
	const double *xpos, *xmat, *vel;
	int representation, apply_forces;
	int32_t *flags;          // [0] capacity overflow bits (16: a contribution was not finite or out of the accumulators'
	                         // range), [1] broadphase queue overflow
	hcs_face *faces;         // optional per-face dump
	int32_t *face_count;
	int max_faces;
	double *face_verts;      // [max_faces][24] world vertices of the dumped faces, or NULL (hcs_config.face_vertices)
	TactileTri *tri_pool;
	int32_t *tri_count;
	int max_tris;
	double *tri_vd;            // [max_tris][9] the same triangles' world vertices in double, or NULL (only taxel
	                           // sensors need them: their sample lattice is evaluated in double)
	uint2 *tri_elem;           // [max_tris] (elemM, elemN) of the polygon a triangle fans, or NULL (only area-importance
	                           // taxel sensors need them: their triangle order is (pair, elemM, elemN, fan triangle))
	hcs_pair_result *pair_out; // [n_env][n_pairs]
	double *geom_wrench;       // [n_env][n_geoms][6]
	// end-to-end path without sensors: device addresses of the context's mapped pinned mirrors (else NULL); the
	// finalize kernel writes the wrenches and the flags there itself, so no copy follows the kernels
	double *geom_wrench_host;
	int32_t *flags_host;
	// The step counters exist twice.  A step works on one set and its finalize kernel, the last kernel that every step has,
	// zeroes the OTHER set for the step after it (n_zero ints; NULL: the host clears the set with a memset before the step),
	// so a step is kernels only: no memset node in front of the chain (C1 x 4096: ~2 us of a 94 us step, more of a
	// single-environment step).  Getters read the set of the last step, which stays intact until the step after it ends.
	int32_t *zero_next;
	int n_zero;
};

struct SensorDev {
	int geom, cx, cy, S, window;
	float sigma;
	double resolution;
	double size[3];
	float rmean, rS;
	const float *weights; // [S*S] window weights
	float *image;         // [n_env][cx*cy]
	int32_t *bin_count;   // [n_env * cx*cy] (triangle, taxel) overlaps per taxel
	int32_t *bin_offset;  // [n_env * cx*cy + 1] exclusive scan of bin_count
	int32_t *bin_cursor;  // [n_env * cx*cy] fill cursors
	int32_t *bin_items;   // [items_cap] triangle ids, bins back to back
	int32_t *scan_tmp;    // tile sums of the scan
	int32_t *raster_counter; // next group of 32 taxels of the raster kernel's persistent warps (zeroed by the clear kernel)
	int items_cap;
};

// CurvedSensor (SENS/src/curved_sensor.cpp): rays from surface sample points along the inward normal, grouped into
// taxels with weights.  The rays are static in the geom frame, so they sit in a static uniform grid (only cells
// that hold rays get an id); every step the contact-surface triangles are binned into the cells whose box, grown
// by the ray length, they overlap, and one warp per (env, cell) finds each ray's nearest hit in its cell's bin.
struct CurvedDev {
	int geom, n_rays, n_taxels, n_cells;   // n_cells = cells that hold rays
	double include_margin;
	double origin[3], cell;                // grid in the geom frame
	int dims[3];
	const double *ray_pos, *ray_nrm;       // [n_rays][3] geom frame
	const int32_t *cell_lookup;            // [dims0*dims1*dims2] -> cell id or -1
	const int32_t *cell_ray_off, *cell_rays; // CSR: rays of each cell
	const int32_t *taxel_off, *taxel_ray;  // CSR: rays of each taxel (load(): surface_idx)
	const double *taxel_w;                 //      and their weights (surface_weight)
	double *raw;                           // [n_env][n_rays] e_MN at the accepted hit, else 0
	float *values;                         // [n_env][n_taxels]
	int32_t *bin_count, *bin_offset, *bin_cursor, *bin_items, *scan_tmp; // per (env, cell) triangle bins
	int items_cap;
};

// TaxelSensor (SENS/src/taxel_sensor.cpp, sample_method "default"): a barycentric lattice of samples on every
// contact-surface triangle, taxel values from the samples within include_margin.  Triangles are binned per
// (env, taxel); one warp per (env, taxel) walks its bin in a canonical order.
struct TaxelDev {
	int geom, n_taxels, method, visualize; // method: 0 closest 1 weighted 2 mean 3 squared (:79-91)
	double include_margin, sample_resolution;
	const double *taxel_pos; // [n_taxels][3] geom frame
	float *values;           // [n_env][n_taxels]; persistent: taxels without a sample in range keep their value
	int32_t *env_tris;       // [n_env] triangles of this sensor's contact surfaces (0: the message is zeroed)
	int32_t *bin_count, *bin_offset, *bin_cursor, *bin_items, *scan_tmp; // per (env, taxel) triangle bins
	int items_cap;
	// sample_method AREA_IMPORTANCE (taxel_sensor.cpp:211-254): per-environment triangle lists and sample records
	int sample_method;                            // 0 DEFAULT, 1 AREA_IMPORTANCE
	int32_t *env_offset, *env_cursor, *env_items; // [n_env + 1], [n_env], [max_tris]
	int32_t *n_samples;                           // [n_env]
	double *samples;                              // [n_env][max_samples][4]: world point, pressure
	int max_samples;
};

#ifdef __CUDACC__
struct GeomBounds {
	double c[3], r;
	float lo[3], hi[3];
};
// bounding sphere and root box of a geom in one environment
__device__ __forceinline__ GeomBounds geom_bounds(const GeomDev &g, int env)
{
	GeomBounds b;
	if (g.env_bounds) {
		const double *p = g.env_bounds + (size_t)env * ENV_BOUNDS;
#pragma unroll
		for (int a = 0; a < 3; ++a)
			b.c[a] = p[a], b.lo[a] = (float)p[4 + a], b.hi[a] = (float)p[7 + a];
		b.r = p[3];
	} else {
#pragma unroll
		for (int a = 0; a < 3; ++a)
			b.c[a] = g.bound_c[a], b.lo[a] = g.root_lo[a], b.hi[a] = g.root_hi[a];
		b.r = g.bound_r;
	}
	return b;
}
#endif

// ---- programmatic dependent launch (sm_90+) -------------------------------------------------------
// The step is a chain of short kernels (C1: 50 + 39 + 14 us).  A kernel launched with launch_chained may become
// resident while its predecessor in the stream is still draining: the predecessor's CTAs call pdl_release() on entry,
// the successor's CTAs fill the SMs as those CTAs exit and block in pdl_wait() until the predecessor grid has
// completed and its writes are visible.  Nothing before pdl_wait() may touch data the predecessor produces.
// Launched the ordinary way (HCS_NO_PDL=1, or with stage events in between) both calls are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Opt a kernel in to more than the default 48 KB of dynamic shared memory, once per (kernel, device) and size (contexts may
// live on several devices and are stepped from several host threads; the call costs a microsecond or two of host time,
// which the end-to-end path of small steps notices).  Sizes up to 48 KB need no opt-in and MUST NOT get one: the attribute
// is a limit, so setting it to a small launch's size makes a later, larger launch of the same kernel fail with "invalid
// argument".  (Until the third session of round 2 the bookkeeping was one array per function-pointer TYPE: kernels of equal
// signature shared it, e.g. the two finalize kernels, and a scene with three pairs followed by one with four could hit
// exactly that: found by the multi-device test on a 2-GPU box.)
template <class K>
static inline void ensure_dynamic_smem(K kernel, int bytes)
{
	if (bytes <= 48 * 1024)
		return;
	static std::mutex mu;
	static std::map<std::pair<const void *, int>, int> done;
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lock(mu);
	int &have = done[std::make_pair(reinterpret_cast<const void *>(kernel), dev)];
	if (have < bytes && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess)
		have = bytes;
}

template <class... KArgs, class... Args>
static inline void launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
	static const bool use_pdl = getenv("HCS_NO_PDL") == nullptr;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim          = grid;
	cfg.blockDim         = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream           = s;
	cudaLaunchAttribute at[1];
	at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs    = at;
	cfg.numAttrs = use_pdl ? 1 : 0;
	cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- launchers (definitions in the .cu files) ---------------------------------------------------
void launch_build_tets(const GeomDev &g, cudaStream_t s);
void launch_build_tris(const GeomDev &g, cudaStream_t s);
// per-environment vertices (+ pressures) of a sphere / ellipsoid from the unit mesh and [n_env][3] sizes (device arrays)
void launch_sphere_env_verts(const double *unit, int n_verts, int vol_offset, const double *sizes, int n_env, int is_sphere, double E,
                             double *verts, double *pressure, cudaStream_t s);
// K2: LBVH of a soft geom on the GPU (g.nodes receives n_elems-1 records); glo/ghi = centroid bounds
size_t lbvh_scratch_bytes(int n);
void launch_build_lbvh(const GeomDev &g, const double glo[3], const double ghi[3], void *scratch, cudaStream_t s);

// K0: the refined-octahedron unit sphere on the GPU (kernels_meshgen.cu), bit-identical to mesh_host.cpp
int unit_sphere_max_gpu_level();
void unit_sphere_counts(int level, int *n_verts, int *n_tri);
size_t unit_sphere_scratch_bytes(int level);
void launch_unit_sphere(int level, double *verts, int32_t *tri, void *scratch, int32_t *bad_flag, cudaStream_t s);
void launch_sphere_elems(const int32_t *tri, int n_tri, int soft, int32_t *elems, cudaStream_t s);
// load-time helpers shared by the builders: in-place bitonic sort of n_pad (power of two) 64-bit keys; exclusive scan of n
// int32 values (tile_tmp: n / 1024 + 2 ints, total: one int)
void launch_bitonic_sort_u64(unsigned long long *keys, int n_pad, cudaStream_t s);
void launch_exclusive_scan_i32(const int32_t *in, int32_t *out, int32_t *tile_tmp, int n, int32_t *total, cudaStream_t s);

void launch_broadphase(const PairDesc &P, const StepIO &io, cudaStream_t s);
// chained: the kernel directly behind it in the stream is the one whose output it consumes (launch_chained above)
void launch_narrowphase(const PairDesc &P, const StepIO &io, cudaStream_t s, bool chained);
// pair results and per-geom wrenches from the exact accumulators; returns the number of kernels launched
int launch_finalize(const PairDesc *d_pairs, const StepIO &io, cudaStream_t s, bool chained);

// clear, count, scan, fill, rasterise for all sensors (host copy + device copy of the records); returns the number
// of kernels launched
int launch_tactile(const SensorDev *sensors, const SensorDev *d_sensors, int n_sensors, const StepIO &io,
                   const PairDesc *d_pairs, cudaStream_t s);
int launch_curved(const CurvedDev &cd, const StepIO &io, const PairDesc *d_pairs, cudaStream_t s);
int launch_taxel(const TaxelDev &td, const StepIO &io, const PairDesc *d_pairs, cudaStream_t s);

} // namespace hcs
