// libhcs_b200 — context, resident-geometry management and the C ABI (include/hcs.h).
//
// Host-side mirror of MujocoContactSurfacesPlugin's configuration/dispatch logic
// (mujoco_contact_surfaces_plugin.cpp:208-318, 571-813) for a BATCH of independent environments on one
// GPU.  There is no CPU compute path: every entry point that needs a result launches CUDA kernels and
// fails with HCS_E_CUDA when no device is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "hcs_internal.h"
#include "mesh_host.h"

namespace hcs {

static thread_local std::string g_create_error;
constexpr long ITEM_LIMIT = 1L << 26; // broadphase queue items address 2^26 nodes / elements (kernels_broadphase.cu)

#define CK(call)                                                                                        \
	do {                                                                                                \
		cudaError_t e_ = (call);                                                                        \
		if (e_ != cudaSuccess) {                                                                        \
			char buf_[512];                                                                             \
			snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
			         __LINE__);                                                                         \
			throw std::runtime_error(buf_);                                                             \
		}                                                                                               \
	} while (0)

struct GeomHost {
	int mj_type = 0;
	double size[3] = { 0, 0, 0 };
	double props[5] = { 0, 0, 0, 0, 0 };
	HostMesh mesh;
	GeomDev dev{};
	bool custom = false; // added through hcs_add_soft_mesh / hcs_add_rigid_mesh
	std::vector<double> env_sizes; // [n_envs][3] per-environment sizes (hcs_set_env_sizes), empty: one size for all
	double base_size[3] = { 0, 0, 0 }; // the one size the geom had before per-environment sizes were set
	bool uploaded = false; // the device records (g.dev, g.allocs) are those of the current mesh / sizes: hcs_finalize after a
	                       // new geom pair or a sensor change keeps them (no mesh generation, no LBVH build)
	std::vector<void *> allocs;
	double E() const { return mesh.soft ? props[0] : std::numeric_limits<double>::infinity(); }
	double dissipation() const { return mesh.soft ? props[1] : 1.0; } // ContactProperties ctor, plugin.h:197-198
};

struct SensorHost {
	int geom;
	double resolution;
	int S, window;
	float sigma;
	int cx, cy;
	SensorDev dev{};
	std::vector<float> weights;
	float *h_image = nullptr; // pinned [n_env][cx*cy]
};

// CurvedSensor::load (SENS/src/curved_sensor.cpp:111-380) with caller-supplied surface samples
struct CurvedHost {
	int geom;
	double include_margin;
	std::vector<double> taxel_pos, taxel_nrm;   // [n][3]; zero normal = no 45 degree test
	std::vector<double> ray_pos, ray_nrm;       // "close" samples in first-use order (:352-356)
	std::vector<int32_t> taxel_off, taxel_ray;  // surface_idx as CSR
	std::vector<double> taxel_w;                // surface_weight
	int n_taxels() const { return (int)taxel_pos.size() / 3; }
	int n_rays() const { return (int)ray_pos.size() / 3; }
	CurvedDev dev{};
	float *h_values = nullptr; // pinned [n_env][n_taxels]
};

// TaxelSensor::load (SENS/src/taxel_sensor.cpp:45-156), sample_method "default"
struct TaxelHost {
	int geom, method, visualize;
	int sample_method = 0; // 0 default, 1 area_importance
	double include_margin, sample_resolution;
	std::vector<double> taxel_pos; // [n][3] geom frame
	int n_taxels() const { return (int)taxel_pos.size() / 3; }
	TaxelDev dev{};
	float *h_values = nullptr; // pinned [n_env][n_taxels]
};

struct EmittedCache {
	int64_t step = -1;
	std::vector<int> offset;       // [n_envs + 1]
	std::vector<int32_t> triples;  // (elemM, elemN, nverts)
};

} // namespace hcs

using namespace hcs;

struct hcs_ctx {
	hcs_config cfg{};
	std::string err;
	cudaStream_t stream = nullptr;
	bool own_stream     = false;
	bool finalized      = false;
	std::vector<GeomHost> geoms;
	std::vector<std::pair<int, int>> pairs;
	std::vector<PairDesc> pair_desc;
	std::vector<SensorHost> sensors;
	std::vector<CurvedHost> curved;
	std::vector<TaxelHost> taxel;
	std::vector<void *> step_allocs;
	PairDesc *d_pairs = nullptr;
	std::vector<SensorDev> sensor_dev; // device records of all sensors, host copy + device copy
	SensorDev *d_sensors = nullptr;
	int32_t *d_counters = nullptr; // flags[4], face count, tactile triangle count, then PAIR_COUNTERS ints per pair; TWO
	size_t n_counters   = 0;       // sets of n_counters ints: a step uses one, its finalize kernel zeroes the other
	bool set_clean[2]   = { false, false }; // the set is all zero
	int cur_set         = 0;                // the set of the last step (what the getters read)
	StepIO io{};
	double *d_xpos = nullptr, *d_xmat = nullptr, *d_vel = nullptr; // staging for the host entry point (one block: xpos | xmat | vel)
	double *h_in = nullptr;                                        // pinned copy of small inputs (CUDA-graph path of hcs_step)
	// hcs_step of small batches replays a captured graph (one H2D copy, the kernels, the result copies): [with_sensors]
	cudaGraphExec_t step_graph[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } }; // [with_sensors][counter set]
	int64_t graph_kernels[2][2]      = { { 0, 0 }, { 0, 0 } };
	int steps_since_finalize      = 0;
	hcs_pair_result *h_pair = nullptr;                             // pinned mirrors
	double *h_wrench        = nullptr;
	int32_t *h_flags        = nullptr;
	double *dh_wrench       = nullptr; // device addresses of h_wrench / h_flags (zero-copy end-to-end path)
	int32_t *dh_flags       = nullptr;
	int64_t kernels_last_step = 0;
	bool profiling = false;
	static constexpr int N_AUX = 4;
	cudaStream_t aux[N_AUX]{};
	cudaEvent_t ev_join[N_AUX]{}, ev_fork = nullptr;
	cudaEvent_t ev[8]{};
	float stage_ms[7]{};
	bool results_on_host = false, pairs_on_host = false, sensors_on_host = false, last_with_sensors = false;
	// asynchronous end-to-end pipeline (hcs_step_async / hcs_wait): two slots of input staging and outputs, a copy-in
	// and a copy-out stream, so that the H2D of step i+1 and the D2H of step i-1 overlap the kernels of step i
	static constexpr int DEPTH = 2;
	struct Slot {
		double *d_xpos = nullptr, *d_xmat = nullptr, *d_vel = nullptr, *d_wrench = nullptr;
		int32_t *h_flags = nullptr, *dh_flags = nullptr; // pinned + mapped: the last kernel of a step writes them
		cudaEvent_t ev_in = nullptr, ev_free = nullptr, ev_kdone = nullptr, ev_done = nullptr;
		int64_t ticket = -1; // step that occupies the slot
		bool waited = true;
		int status = HCS_OK;
	} slot[DEPTH];
	cudaStream_t copy_in = nullptr, copy_out = nullptr;
	int64_t next_ticket = 0;
	int64_t step_counter = 0;                 // steps launched so far
	std::vector<EmittedCache> emitted_cache;  // hcs_get_emitted: per pair, the last step's list indexed by environment
};

namespace hcs {

template <class T>
static T *dalloc(std::vector<void *> &bag, size_t n)
{
	void *p = nullptr;
	CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
	bag.push_back(p);
	return (T *)p;
}

static void free_bag(std::vector<void *> &bag)
{
	for (void *p : bag)
		cudaFree(p);
	bag.clear();
}

// ---- LBVH over the tets of a soft geom: Morton codes of element centroids, Karras radix tree,
// children boxes stored in the parent (float, rounded outward).  Built once per geom in its own frame;
// meshes are rigid so the tree is never rebuilt in the step.
static inline uint32_t expand_bits(uint32_t v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
static inline float down(double x)
{
	float f = (float)x;
	return (double)f > x ? std::nextafterf(f, -INFINITY) : f;
}
static inline float up(double x)
{
	float f = (float)x;
	return (double)f < x ? std::nextafterf(f, INFINITY) : f;
}

struct BoxD {
	double lo[3], hi[3];
};

static std::vector<BvhNode> build_lbvh(const HostMesh &m)
{
	const int n = m.n_elems();
	std::vector<BoxD> leaf(n);
	std::vector<double> cen(3 * (size_t)n);
	double glo[3] = { 1e300, 1e300, 1e300 }, ghi[3] = { -1e300, -1e300, -1e300 };
	for (int t = 0; t < n; ++t) {
		BoxD b{ { 1e300, 1e300, 1e300 }, { -1e300, -1e300, -1e300 } };
		double c[3] = { 0, 0, 0 };
		for (int k = 0; k < 4; ++k) {
			const double *p = &m.verts[3 * (size_t)m.elems[4 * (size_t)t + k]];
			for (int a = 0; a < 3; ++a) {
				b.lo[a] = std::min(b.lo[a], p[a]);
				b.hi[a] = std::max(b.hi[a], p[a]);
				c[a] += 0.25 * p[a];
			}
		}
		leaf[t] = b;
		for (int a = 0; a < 3; ++a) {
			cen[3 * (size_t)t + a] = c[a];
			glo[a] = std::min(glo[a], c[a]);
			ghi[a] = std::max(ghi[a], c[a]);
		}
	}
	std::vector<std::pair<uint64_t, int>> keys(n);
	for (int t = 0; t < n; ++t) {
		uint32_t q[3];
		for (int a = 0; a < 3; ++a) {
			double ext = ghi[a] - glo[a];
			double u   = ext > 0 ? (cen[3 * (size_t)t + a] - glo[a]) / ext : 0.0;
			q[a]       = (uint32_t)std::min(1023.0, std::max(0.0, u * 1024.0));
		}
		uint32_t code = expand_bits(q[0]) * 4 + expand_bits(q[1]) * 2 + expand_bits(q[2]);
		keys[t]       = { ((uint64_t)code << 32) | (uint32_t)t, t };
	}
	std::sort(keys.begin(), keys.end());
	std::vector<BvhNode> nodes(std::max(1, n - 1));
	auto set_box = [&](BvhNode &nd, bool left, const BoxD &b) {
		float *lo = left ? nd.llo : nd.rlo, *hi = left ? nd.lhi : nd.rhi;
		for (int a = 0; a < 3; ++a) {
			lo[a] = down(b.lo[a]);
			hi[a] = up(b.hi[a]);
		}
	};
	if (n == 1) {
		BvhNode &nd = nodes[0];
		set_box(nd, true, leaf[keys[0].second]);
		nd.left = ~keys[0].second;
		for (int a = 0; a < 3; ++a)
			nd.rlo[a] = 3e38f, nd.rhi[a] = -3e38f;
		nd.right  = ~keys[0].second;
		nd.pad[0] = nd.pad[1] = 0;
		return nodes;
	}
	// Karras 2012: common-prefix length of the (unique) 64-bit keys
	auto delta = [&](int i, int j) -> int {
		if (j < 0 || j >= n)
			return -1;
		return __builtin_clzll(keys[i].first ^ keys[j].first);
	};
	std::vector<int> left(n - 1), right(n - 1); // >=0 internal, <0 ~leaf slot
	for (int i = 0; i < n - 1; ++i) {
		int d     = delta(i, i + 1) - delta(i, i - 1) >= 0 ? 1 : -1;
		int dmin  = delta(i, i - d);
		int lmax  = 2;
		while (delta(i, i + lmax * d) > dmin)
			lmax *= 2;
		int l = 0;
		for (int t = lmax / 2; t >= 1; t /= 2)
			if (delta(i, i + (l + t) * d) > dmin)
				l += t;
		int j     = i + l * d;
		int dnode = delta(i, j);
		int s     = 0;
		for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
			if (delta(i, i + (s + t) * d) > dnode)
				s += t;
			if (t == 1)
				break;
		}
		int gamma = i + s * d + std::min(d, 0);
		left[i]   = std::min(i, j) == gamma ? ~gamma : gamma;
		right[i]  = std::max(i, j) == gamma + 1 ? ~(gamma + 1) : gamma + 1;
	}
	// bottom-up boxes with an explicit stack (post-order from the root = node 0)
	std::vector<BoxD> nbox(n - 1);
	std::vector<char> done(n - 1, 0);
	std::vector<int> stack = { 0 };
	auto child_box = [&](int c) -> const BoxD & { return c < 0 ? leaf[keys[~c].second] : nbox[c]; };
	while (!stack.empty()) {
		int i = stack.back();
		bool ready = true;
		for (int c : { left[i], right[i] })
			if (c >= 0 && !done[c]) {
				stack.push_back(c);
				ready = false;
			}
		if (!ready)
			continue;
		stack.pop_back();
		const BoxD &a = child_box(left[i]), &b = child_box(right[i]);
		for (int k = 0; k < 3; ++k) {
			nbox[i].lo[k] = std::min(a.lo[k], b.lo[k]);
			nbox[i].hi[k] = std::max(a.hi[k], b.hi[k]);
		}
		done[i]     = 1;
		BvhNode &nd = nodes[i];
		set_box(nd, true, a);
		set_box(nd, false, b);
		nd.left   = left[i] < 0 ? ~keys[~left[i]].second : left[i];
		nd.right  = right[i] < 0 ? ~keys[~right[i]].second : right[i];
		nd.pad[0] = nd.pad[1] = 0;
	}
	return nodes;
}

// bounding sphere about the box centre of the vertices
static void bounding_sphere(const double *verts, int nv, double c[3], double &r)
{
	double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
	for (int i = 0; i < nv; ++i)
		for (int a = 0; a < 3; ++a) {
			lo[a] = std::min(lo[a], verts[3 * (size_t)i + a]);
			hi[a] = std::max(hi[a], verts[3 * (size_t)i + a]);
		}
	double r2 = 0;
	for (int a = 0; a < 3; ++a)
		c[a] = 0.5 * (lo[a] + hi[a]);
	for (int i = 0; i < nv; ++i) {
		double s = 0;
		for (int a = 0; a < 3; ++a) {
			double t = verts[3 * (size_t)i + a] - c[a];
			s += t * t;
		}
		r2 = std::max(r2, s);
	}
	r = std::sqrt(r2) * (1 + 1e-12) + 1e-12;
}

// bounds of the tet centroids: they fix the Morton quantisation grid of the LBVH build
static void centroid_bounds(const double *verts, const int32_t *elems, int ne, double glo[3], double ghi[3])
{
	for (int a = 0; a < 3; ++a)
		glo[a] = 1e300, ghi[a] = -1e300;
	for (int t = 0; t < ne; ++t) {
		double cc[3] = { 0, 0, 0 };
		for (int k = 0; k < 4; ++k)
			for (int a = 0; a < 3; ++a)
				cc[a] += 0.25 * verts[3 * (size_t)elems[4 * (size_t)t + k] + a];
		for (int a = 0; a < 3; ++a) {
			glo[a] = std::min(glo[a], cc[a]);
			ghi[a] = std::max(ghi[a], cc[a]);
		}
	}
}

// The unit sphere of a refinement level, generated on the GPU (kernels_meshgen.cu); the arrays go into `bag`
static void device_unit_sphere(hcs_ctx *c, int level, std::vector<void *> &bag, double **d_unit, int32_t **d_tri, int *nv, int *nt)
{
	unit_sphere_counts(level, nv, nt);
	*d_unit = dalloc<double>(bag, (size_t)*nv * 3);
	*d_tri  = dalloc<int32_t>(bag, (size_t)*nt * 3);
	void *scratch = nullptr;
	CK(cudaMalloc(&scratch, unit_sphere_scratch_bytes(level) + sizeof(int32_t)));
	int32_t *d_bad = reinterpret_cast<int32_t *>(static_cast<char *>(scratch) + unit_sphere_scratch_bytes(level));
	launch_unit_sphere(level, *d_unit, *d_tri, scratch, d_bad, c->stream);
	int32_t bad = 0;
	CK(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaGetLastError());
	cudaFree(scratch);
	if (bad)
		throw std::runtime_error("GPU sphere refinement: the boundary is not a closed surface");
}

static bool gpu_meshgen_wanted(const GeomHost &g, int level)
{
	return !g.custom && (g.mj_type == HCS_GEOM_SPHERE || g.mj_type == HCS_GEOM_ELLIPSOID) && level <= unit_sphere_max_gpu_level() &&
	       getenv("HCS_MESHGEN_HOST") == nullptr;
}

// Sphere and ellipsoid geoms: topology, vertices and pressures generated on the GPU (K0), mirrored into g.mesh for the
// host-side bookkeeping (bounds, pool sizes, getters).  hcs_add_geom has validated the geom with the host generator and
// left its mesh in g.mesh: with HCS_MESHGEN_CHECK set the two are compared bit for bit.
static void gpu_generate_sphere_like(hcs_ctx *c, GeomHost &g)
{
	const int level = sphere_like_level(g.mj_type, g.size, g.props[2]);
	if (!gpu_meshgen_wanted(g, level) || g.mesh.plane)
		return;
	std::vector<void *> tmp;
	try {
		double *d_unit = nullptr;
		int32_t *d_tri = nullptr;
		int nv_vol = 0, nt = 0;
		device_unit_sphere(c, level, tmp, &d_unit, &d_tri, &nv_vol, &nt);
		const bool soft      = g.mesh.soft;
		const int vol_offset = soft ? 0 : 1, nv = nv_vol - vol_offset, per = soft ? 4 : 3;
		double *d_verts  = dalloc<double>(tmp, (size_t)nv * 3);
		double *d_press  = soft ? dalloc<double>(tmp, (size_t)nv) : nullptr;
		int32_t *d_elems = dalloc<int32_t>(tmp, (size_t)nt * per);
		double *d_size   = dalloc<double>(tmp, 3);
		CK(cudaMemcpyAsync(d_size, g.size, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		launch_sphere_env_verts(d_unit, nv, vol_offset, d_size, 1, g.mj_type == HCS_GEOM_SPHERE, g.props[0], d_verts, d_press, c->stream);
		launch_sphere_elems(d_tri, nt, soft ? 1 : 0, d_elems, c->stream);
		HostMesh gm;
		gm.soft = soft;
		gm.verts.resize((size_t)nv * 3), gm.elems.resize((size_t)nt * per), gm.pressure.resize(soft ? nv : 0);
		CK(cudaMemcpyAsync(gm.verts.data(), d_verts, gm.verts.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaMemcpyAsync(gm.elems.data(), d_elems, gm.elems.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
		if (soft)
			CK(cudaMemcpyAsync(gm.pressure.data(), d_press, gm.pressure.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		CK(cudaGetLastError());
		if (getenv("HCS_MESHGEN_CHECK")) {
			auto same = [](const std::vector<double> &a, const std::vector<double> &b) {
				return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
			};
			if (!same(gm.verts, g.mesh.verts) || gm.elems != g.mesh.elems || !same(gm.pressure, g.mesh.pressure))
				throw std::runtime_error("HCS_MESHGEN_CHECK: the GPU-generated sphere mesh differs from the host generator's");
		}
		g.mesh = std::move(gm);
	} catch (...) {
		free_bag(tmp);
		throw;
	}
	free_bag(tmp);
}

// Per-environment sizes (hcs_set_env_sizes): ONE topology (g.mesh, the mesh of environment 0), every
// environment its own vertices, pressures, element records and LBVH, env-major.  Sphere / ellipsoid vertices and pressures
// are generated on the GPU from the unit mesh; other shapes come from the host generator, environment by environment.
// Fields, element records (K1) and the LBVH (K2) are built on the GPU for every environment.
static void upload_geom_per_env(hcs_ctx *c, GeomHost &g)
{
	const int n_env   = c->cfg.n_envs;
	const HostMesh &m = g.mesh;
	if (m.plane || g.custom || (int)g.env_sizes.size() != 3 * n_env)
		throw std::runtime_error("per-environment sizes: needs a generated (non-plane) geom and n_envs x 3 sizes");
	GeomDev d{};
	d.kind    = m.soft ? 1 : 0;
	d.n_verts = m.n_verts();
	d.n_elems = m.n_elems();
	const int nv = d.n_verts, ne = d.n_elems, per = m.soft ? 4 : 3;
	if (m.soft && ne < 2)
		throw std::runtime_error("per-environment sizes: the soft geom needs at least two tets");
	d.elems = dalloc<int32_t>(g.allocs, m.elems.size());
	CK(cudaMemcpyAsync(d.elems, m.elems.data(), m.elems.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
	d.verts = dalloc<double>(g.allocs, (size_t)n_env * nv * 3);
	if (m.soft)
		d.pressure = dalloc<double>(g.allocs, (size_t)n_env * nv);
	std::vector<double> verts((size_t)n_env * nv * 3);
	const bool sphere_like = g.mj_type == HCS_GEOM_SPHERE || g.mj_type == HCS_GEOM_ELLIPSOID;
	if (sphere_like) {
		const int level = sphere_like_level(g.mj_type, &g.env_sizes[0], g.props[2]);
		for (int e = 1; e < n_env; ++e)
			if (sphere_like_level(g.mj_type, &g.env_sizes[3 * (size_t)e], g.props[2]) != level)
				throw std::runtime_error("per-environment sizes: environment " + std::to_string(e) +
				                         " needs another refinement level than environment 0 (scale the resolution hint with the size)");
		const int vol_offset = m.soft ? 0 : 1; // the rigid surface drops the centre vertex
		double *d_unit = nullptr;
		std::vector<double> unit;
		if (gpu_meshgen_wanted(g, level)) { // unit mesh and element list from the GPU generator (K0)
			int32_t *d_tri = nullptr;
			int nv_vol = 0, nt = 0;
			device_unit_sphere(c, level, g.allocs, &d_unit, &d_tri, &nv_vol, &nt);
			if (nv_vol != nv + vol_offset || nt != ne)
				throw std::runtime_error("per-environment sizes: unit sphere and geom mesh disagree");
			launch_sphere_elems(d_tri, nt, m.soft ? 1 : 0, d.elems, c->stream);
		} else {
			unit_sphere_vertices(level, unit);
			if ((int)unit.size() / 3 != nv + vol_offset)
				throw std::runtime_error("per-environment sizes: unit sphere and geom mesh disagree");
			d_unit = dalloc<double>(g.allocs, unit.size());
			CK(cudaMemcpyAsync(d_unit, unit.data(), unit.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		}
		double *d_sizes = dalloc<double>(g.allocs, g.env_sizes.size());
		CK(cudaMemcpyAsync(d_sizes, g.env_sizes.data(), g.env_sizes.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		launch_sphere_env_verts(d_unit, nv, vol_offset, d_sizes, n_env, g.mj_type == HCS_GEOM_SPHERE, g.props[0], d.verts, d.pressure,
		                        c->stream);
		CK(cudaMemcpyAsync(verts.data(), d.verts, verts.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream)); // (the host needs the vertices for the bounds below; `unit` is a local)
	} else {
		std::vector<double> pressure(m.soft ? (size_t)n_env * nv : 0);
		for (int e = 0; e < n_env; ++e) {
			HostMesh me;
			std::string err;
			if (!build_geom_mesh(g.mj_type, &g.env_sizes[3 * (size_t)e], nullptr, 0, nullptr, 0, g.props, me, err))
				throw std::runtime_error("per-environment sizes: environment " + std::to_string(e) + ": " + err);
			if (me.elems != m.elems || me.n_verts() != nv)
				throw std::runtime_error("per-environment sizes: environment " + std::to_string(e) +
				                         " leads to another mesh topology than environment 0");
			std::copy(me.verts.begin(), me.verts.end(), verts.begin() + (size_t)e * nv * 3);
			if (m.soft)
				std::copy(me.pressure.begin(), me.pressure.end(), pressure.begin() + (size_t)e * nv);
		}
		CK(cudaMemcpyAsync(d.verts, verts.data(), verts.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		if (m.soft)
			CK(cudaMemcpyAsync(d.pressure, pressure.data(), pressure.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CK(cudaStreamSynchronize(c->stream)); // `pressure` is a local
	}
	d.env_stride       = ne;
	d.env_stride_verts = nv;
	d.env_stride_nodes = m.soft ? ne - 1 : 0;
	std::vector<double> bounds((size_t)n_env * ENV_BOUNDS, 0.0);
	if (m.soft) {
		d.tet_geom     = dalloc<TetGeom>(g.allocs, (size_t)n_env * ne);
		d.tet_field    = dalloc<TetField>(g.allocs, (size_t)n_env * ne);
		d.tet_leaf32   = dalloc<TetLeaf32>(g.allocs, (size_t)n_env * ne);
		d.tet_leafss32 = dalloc<TetLeafSS32>(g.allocs, (size_t)n_env * ne);
		d.tet_box32    = dalloc<TetBox32>(g.allocs, (size_t)n_env * ne);
		d.nodes        = dalloc<BvhNode>(g.allocs, (size_t)n_env * (ne - 1));
	} else {
		d.tris = dalloc<TriRec>(g.allocs, (size_t)n_env * ne);
	}
	void *scratch = nullptr;
	if (m.soft)
		CK(cudaMalloc(&scratch, lbvh_scratch_bytes(ne)));
	for (int e = 0; e < n_env; ++e) {
		GeomDev v = d; // environment e's slice of every array
		v.verts   = d.verts + (size_t)e * nv * 3;
		const double *hv = verts.data() + (size_t)e * nv * 3;
		bounding_sphere(hv, nv, &bounds[(size_t)e * ENV_BOUNDS], bounds[(size_t)e * ENV_BOUNDS + 3]);
		if (m.soft) {
			v.pressure     = d.pressure + (size_t)e * nv;
			v.tet_geom     = d.tet_geom + (size_t)e * ne;
			v.tet_field    = d.tet_field + (size_t)e * ne;
			v.tet_leaf32   = d.tet_leaf32 + (size_t)e * ne;
			v.tet_leafss32 = d.tet_leafss32 + (size_t)e * ne;
			v.tet_box32    = d.tet_box32 + (size_t)e * ne;
			v.nodes        = d.nodes + (size_t)e * (ne - 1);
			double glo[3], ghi[3];
			centroid_bounds(hv, m.elems.data(), ne, glo, ghi);
			launch_build_lbvh(v, glo, ghi, scratch, c->stream);
			launch_build_tets(v, c->stream);
		} else {
			v.tris = d.tris + (size_t)e * ne;
			launch_build_tris(v, c->stream);
		}
	}
	if (m.soft) { // root boxes: node 0 of every environment
		std::vector<BvhNode> roots(n_env);
		CK(cudaMemcpy2DAsync(roots.data(), sizeof(BvhNode), d.nodes, (size_t)(ne - 1) * sizeof(BvhNode), sizeof(BvhNode), n_env,
		                     cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		cudaFree(scratch);
		for (int e = 0; e < n_env; ++e)
			for (int a = 0; a < 3; ++a) {
				bounds[(size_t)e * ENV_BOUNDS + 4 + a] = std::min(roots[e].llo[a], roots[e].rlo[a]);
				bounds[(size_t)e * ENV_BOUNDS + 7 + a] = std::max(roots[e].lhi[a], roots[e].rhi[a]);
			}
	}
	CK(cudaGetLastError());
	double *d_bounds = dalloc<double>(g.allocs, bounds.size());
	CK(cudaMemcpyAsync(d_bounds, bounds.data(), bounds.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream)); // `bounds` is a local
	d.env_bounds = d_bounds;
	// the members that describe one geometry: environment 0; bound_r: the largest (pool sizing, accumulator scale)
	for (int a = 0; a < 3; ++a) {
		d.bound_c[a] = bounds[a];
		d.root_lo[a] = (float)bounds[4 + a], d.root_hi[a] = (float)bounds[7 + a];
	}
	d.bound_r = 0;
	for (int e = 0; e < n_env; ++e)
		d.bound_r = std::max(d.bound_r, bounds[(size_t)e * ENV_BOUNDS + 3]);
	g.dev = d;
}

static void upload_geom(hcs_ctx *c, GeomHost &g)
{
	free_bag(g.allocs);
	if (!g.env_sizes.empty()) {
		upload_geom_per_env(c, g);
		return;
	}
	if (!g.custom && (g.mj_type == HCS_GEOM_SPHERE || g.mj_type == HCS_GEOM_ELLIPSOID))
		gpu_generate_sphere_like(c, g);
	GeomDev d{};
	const HostMesh &m = g.mesh;
	d.kind            = m.plane ? 2 : (m.soft ? 1 : 0);
	d.n_verts         = m.n_verts();
	d.n_elems         = m.n_elems();
	if (!m.plane) {
		d.verts = dalloc<double>(g.allocs, m.verts.size());
		d.elems = dalloc<int32_t>(g.allocs, m.elems.size());
		CK(cudaMemcpyAsync(d.verts, m.verts.data(), m.verts.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CK(cudaMemcpyAsync(d.elems, m.elems.data(), m.elems.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
		// bounding sphere about the box centre of the vertices
		double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
		for (int i = 0; i < d.n_verts; ++i)
			for (int a = 0; a < 3; ++a) {
				lo[a] = std::min(lo[a], m.verts[3 * (size_t)i + a]);
				hi[a] = std::max(hi[a], m.verts[3 * (size_t)i + a]);
			}
		double r2 = 0;
		for (int a = 0; a < 3; ++a)
			d.bound_c[a] = 0.5 * (lo[a] + hi[a]);
		for (int i = 0; i < d.n_verts; ++i) {
			double s = 0;
			for (int a = 0; a < 3; ++a) {
				double t = m.verts[3 * (size_t)i + a] - d.bound_c[a];
				s += t * t;
			}
			r2 = std::max(r2, s);
		}
		d.bound_r = std::sqrt(r2) * (1 + 1e-12) + 1e-12;
		if (m.soft) {
			d.pressure = dalloc<double>(g.allocs, m.pressure.size());
			CK(cudaMemcpyAsync(d.pressure, m.pressure.data(), m.pressure.size() * sizeof(double), cudaMemcpyHostToDevice,
			                   c->stream));
			d.tet_geom  = dalloc<TetGeom>(g.allocs, d.n_elems);
			d.tet_field = dalloc<TetField>(g.allocs, d.n_elems);
			d.tet_leaf32 = dalloc<TetLeaf32>(g.allocs, d.n_elems);
			d.tet_leafss32 = dalloc<TetLeafSS32>(g.allocs, d.n_elems);
			d.tet_box32 = dalloc<TetBox32>(g.allocs, d.n_elems);
			d.nodes = dalloc<BvhNode>(g.allocs, std::max(1, d.n_elems - 1));
			BvhNode root;
			if (d.n_elems < 2 || std::getenv("HCS_LBVH_HOST")) { // host builder: single-tet trees and cross-checks
				std::vector<BvhNode> nodes = build_lbvh(m);
				CK(cudaMemcpyAsync(d.nodes, nodes.data(), nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice, c->stream));
				CK(cudaStreamSynchronize(c->stream)); // `nodes` is a local
				root = nodes[0];
			} else { // K2: Morton codes, sort, Karras tree and refit on the GPU
				double glo[3] = { 1e300, 1e300, 1e300 }, ghi[3] = { -1e300, -1e300, -1e300 };
				for (int t = 0; t < d.n_elems; ++t) { // centroid bounds fix the Morton quantisation grid
					double cc[3] = { 0, 0, 0 };
					for (int k = 0; k < 4; ++k)
						for (int a = 0; a < 3; ++a)
							cc[a] += 0.25 * m.verts[3 * (size_t)m.elems[4 * (size_t)t + k] + a];
					for (int a = 0; a < 3; ++a) {
						glo[a] = std::min(glo[a], cc[a]);
						ghi[a] = std::max(ghi[a], cc[a]);
					}
				}
				void *scratch = nullptr;
				CK(cudaMalloc(&scratch, lbvh_scratch_bytes(d.n_elems)));
				launch_build_lbvh(d, glo, ghi, scratch, c->stream);
				CK(cudaMemcpyAsync(&root, d.nodes, sizeof(BvhNode), cudaMemcpyDeviceToHost, c->stream));
				CK(cudaStreamSynchronize(c->stream));
				CK(cudaGetLastError());
				cudaFree(scratch);
			}
			for (int a = 0; a < 3; ++a) {
				d.root_lo[a] = std::min(root.llo[a], root.rlo[a]);
				d.root_hi[a] = std::max(root.lhi[a], root.rhi[a]);
			}
			if (d.n_elems >= SPLIT_MIN_TREE) { // subtree roots at depth SPLIT_DEPTH for the flat traversal
				std::vector<BvhNode> nodes((size_t)d.n_elems - 1);
				CK(cudaMemcpyAsync(nodes.data(), d.nodes, nodes.size() * sizeof(BvhNode), cudaMemcpyDeviceToHost, c->stream));
				CK(cudaStreamSynchronize(c->stream));
				std::vector<int32_t> level{ 0 }, split;
				for (int depth = 0; depth < SPLIT_DEPTH; ++depth) {
					std::vector<int32_t> next;
					for (int32_t n : level) {
						for (int32_t ch : { nodes[n].left, nodes[n].right }) {
							if (ch < 0)
								split.push_back(ch); // a leaf above the split depth is a subtree of its own
							else
								next.push_back(ch);
						}
					}
					level.swap(next);
				}
				split.insert(split.end(), level.begin(), level.end());
				int32_t *d_split = dalloc<int32_t>(g.allocs, split.size());
				CK(cudaMemcpyAsync(d_split, split.data(), split.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
				CK(cudaStreamSynchronize(c->stream)); // `split` is a local
				d.split_nodes = d_split;
				d.n_split     = (int)split.size();
			}
			launch_build_tets(d, c->stream);
		} else {
			d.tris = dalloc<TriRec>(g.allocs, d.n_elems);
			launch_build_tris(d, c->stream);
		}
		CK(cudaGetLastError());
	} else {
		d.bound_r = 1e300;
	}
	g.dev = d;
}

// plugin.cpp:128-159
static double combined_dissipation(const GeomHost &a, const GeomHost &b)
{
	const double inf = std::numeric_limits<double>::infinity();
	double EA = a.E(), EB = b.E(), dA = a.dissipation(), dB = b.dissipation();
	double Es = EA == inf ? EB : (EB == inf ? EA : EA * EB / (EA + EB));
	if (Es == inf)
		return 0.5 * (dA + dB);
	double d = 0;
	if (EA != inf)
		d += Es / EA * dA;
	if (EB != inf)
		d += Es / EB * dB;
	return d;
}
// drake CalcContactFrictionFromSurfaceProperties, dynamic coefficient (plugin.cpp:427-433)
static double combined_mu(const GeomHost &a, const GeomHost &b)
{
	double den = a.props[4] + b.props[4];
	return den == 0 ? 0.0 : 2 * a.props[4] * b.props[4] / den;
}

static bool is_sensor_geom(const hcs_ctx *c, int g)
{
	for (const SensorHost &s : c->sensors)
		if (s.geom == g)
			return true;
	for (const CurvedHost &s : c->curved)
		if (s.geom == g)
			return true;
	for (const TaxelHost &s : c->taxel)
		if (s.geom == g)
			return true;
	return false;
}

static void build_pairs(hcs_ctx *c)
{
	const int n_env = c->cfg.n_envs;
	c->pair_desc.clear();
	const char *tu_env = getenv("HCS_TARGET_UNITS"); // tuning knob: warps per launch the slicing aims for
	const long target_units = tu_env ? atol(tu_env) : 4096;
	for (size_t pi = 0; pi < c->pairs.size(); ++pi) {
		int g1 = c->pairs[pi].first, g2 = c->pairs[pi].second;
		const GeomHost *c1 = &c->geoms[g1], *c2 = &c->geoms[g2];
		PairDesc P{};
		P.index = (int)pi;
		P.gM    = std::min(g1, g2);
		P.gN    = std::max(g1, g2);
		P.kind  = PAIR_NONE;
		bool s1 = c1->mesh.soft, s2 = c2->mesh.soft;
		if (s1 || s2) { // collision_cb dispatch, plugin.cpp:280-305
			if (s1 && s2) {
				P.kind = PAIR_SOFT_SOFT;
				P.gA = g1, P.gB = g2;
			} else {
				if (!s1) {
					std::swap(c1, c2);
					std::swap(g1, g2);
				}
				P.gA = g1, P.gB = g2;
				P.kind = c2->mesh.plane ? PAIR_SOFT_PLANE : PAIR_SOFT_RIGID;
			}
			const GeomHost &A = c->geoms[P.gA], &B = c->geoms[P.gB];
			P.A = A.dev, P.B = B.dev;
			P.sign        = P.gA == P.gM ? 1.0 : -1.0;
			P.dissipation = combined_dissipation(c->geoms[P.gM], c->geoms[P.gN]);
			P.mu          = combined_mu(c->geoms[P.gM], c->geoms[P.gN]);
			P.n_tree      = A.dev.n_elems;
			P.nq          = P.kind == PAIR_SOFT_PLANE ? A.dev.n_elems : B.dev.n_elems;
			P.emit_tactile = c->cfg.representation == HCS_REP_TRIANGLE && (is_sensor_geom(c, P.gA) || is_sensor_geom(c, P.gB));
			int chunks   = (P.nq + 31) / 32;
			long want    = std::max<long>(1, (target_units + n_env - 1) / n_env);
			int S        = (int)std::min<long>(chunks, want);
			S            = std::max(S, 1);
			P.slice_q    = 32 * ((chunks + S - 1) / S);
			// Small batches against a LARGE tree (the reference's own case is ONE mjData; C5: 131 072-tet pads): a unit is
			// walked by one warp, and the few query elements that touch the other geom are neighbours in the element
			// order, so with 32 queries per unit one or two warps did all the work of an environment.  When the batch
			// cannot fill the GPU with 32-query units the slices shrink, down to a single query element (its tree walk
			// still pops up to 32 nodes per iteration): C5 x 1 env 4.74 -> 2.62 ms.  Small trees keep whole chunks: there
			// the extra units cost more in the finalize than the walk gains (C1 x 1 env 0.028 -> 0.037 ms, C3 x 1 env
			// 0.075 -> 0.34 ms; scripts/sweep_r01j.sh).  HCS_FINE_SLICES=0/1 overrides.
			{
				const char *fs = getenv("HCS_FINE_SLICES");
				const bool fine = fs ? atoi(fs) != 0 : P.n_tree >= 16384;
				if (fine && P.kind != PAIR_SOFT_PLANE && want > chunks)
					P.slice_q = (int)std::max<long>(1, (P.nq + want - 1) / want);
			}
			P.n_slices   = (P.nq + P.slice_q - 1) / P.slice_q;
			if (P.n_slices > (1 << TRI_SLICE_BITS)) { // the slice index is part of the triangle ordering key
				P.slice_q  = (P.nq + (1 << TRI_SLICE_BITS) - 1) >> TRI_SLICE_BITS;
				P.n_slices = (P.nq + P.slice_q - 1) / P.slice_q;
			}
			if (P.nq > (1 << 24) || P.n_tree > (int)ITEM_LIMIT)
				throw std::runtime_error("geom pair too large: at most 2^24 query elements and 2^26 tree elements");
			{
				size_t units = (size_t)n_env * P.n_slices;
				// ONE candidate list per pair for the whole batch (16-byte records + 1 byte each).  Default size: per
				// environment min(nq * n_tree, 16 (nq + n_tree)) candidates, at most 64 M entries;
				// hcs_config.max_candidates_per_slice > 0 sizes it as that many per (env, slice) unit instead,
				// HCS_MAX_TOTAL_CANDIDATES overrides both.  Overflow is reported by hcs_step, never UB.
				// (half-space pairs: the candidates are the tets the plane cuts, at most all of them)
				long per_env = P.kind == PAIR_SOFT_PLANE ? (long)P.nq :
				                                           std::min<long>((long)P.nq * P.n_tree, 16L * ((long)P.nq + P.n_tree));
				long total   = c->cfg.max_candidates_per_slice > 0 ? (long)c->cfg.max_candidates_per_slice * (long)units :
				                                                     std::min<long>(per_env * n_env, 64L << 20);
				if (const char *mt_env = getenv("HCS_MAX_TOTAL_CANDIDATES")) {
					char *end = nullptr;
					long v    = strtol(mt_env, &end, 10);
					if (end == mt_env || *end != 0 || v < 1)
						throw std::runtime_error("HCS_MAX_TOTAL_CANDIDATES must be a positive integer");
					total = v;
				}
				P.contrib_cap = (int)std::max<long>(1024, std::min<long>(total, 1L << 30));
				P.flat        = dalloc<uint4>(c->step_allocs, (size_t)P.contrib_cap);
				P.nverts      = dalloc<uint8_t>(c->step_allocs, (size_t)P.contrib_cap);
				P.accum       = dalloc<int64_t>(c->step_allocs, (size_t)n_env * ACC_WORDS);
				{
					// Power-of-two scale of the exact accumulators (hcs_internal.h): the pair's force scale is modulus x
					// contact area bound (pi R^2 of the smaller geom); torques and area moments about the world origin get
					// 2^10 m of lever arm, and 2^14 on top of that is left for dissipation factors and rough estimates.
					const double inf = std::numeric_limits<double>::infinity();
					double E = std::min(A.E(), B.E());
					if (!(E > 0) || E == inf)
						E = 1.0;
					double R = A.dev.bound_r;
					if (B.dev.bound_r < R && B.dev.bound_r > 0)
						R = B.dev.bound_r;
					const double F0 = std::max(E, 1.0) * M_PI * R * R * 1024.0;
					int k = 0;
					std::frexp(F0, &k); // F0 < 2^k
					k -= 12;
					P.acc_scale   = std::ldexp(1.0, -k);
					P.acc_unscale = std::ldexp(1.0, k);
				}
				// flat broadphase: alive-query records, one per (env, query element) at most (4 GB cap; overflow is reported)
				P.alive = nullptr, P.alive_cap = 0;
				if (P.kind != PAIR_SOFT_PLANE) {
					P.alive_cap = (int)std::min<long>((long)n_env * P.nq, 32L << 20);
					P.alive     = dalloc<float>(c->step_allocs, (size_t)P.alive_cap * ALIVE_WORDS);
				}
				P.counters    = c->d_counters + 6 + PAIR_COUNTERS * pi;
				P.pair_ctx    = dalloc<double>(c->step_allocs, (size_t)n_env * PAIR_CTX_DOUBLES);
				CK(cudaMemsetAsync(P.accum, 0, (size_t)n_env * ACC_WORDS * sizeof(int64_t), c->stream)); // the finalize kernel keeps them zero
			}
		}
		c->pair_desc.push_back(P);
	}
	c->d_pairs = dalloc<PairDesc>(c->step_allocs, c->pair_desc.size());
	CK(cudaMemcpyAsync(c->d_pairs, c->pair_desc.data(), c->pair_desc.size() * sizeof(PairDesc), cudaMemcpyHostToDevice,
	                   c->stream));
}

// window weights of FlatTactileSensor (flat_tactile_sensor.cpp:179-185, 305-314, 353-388), which depend
// only on the sub-sample (i, j): tabulated once on the host with the same float/double mix.
static void build_sensor(hcs_ctx *c, SensorHost &s)
{
	const GeomHost &g = c->geoms[s.geom];
	const int S       = s.S;
	double resolution = s.resolution;
	float di_factor     = resolution / S;
	float sub_halfwidth = di_factor / 2.0f - resolution / 2.0f;
	float rmean         = 1. / (S * S);
	float rS            = resolution / S;
	const float SQRT_2  = 1.41421356237;
	float max_dist      = SQRT_2 * resolution / 2.0f;
	float sigma         = s.sigma;
	float rsigma_squared = 0, rtukey = 0;
	if (s.window == HCS_WINDOW_GAUSS)
		rsigma_squared = 0.5f / (sigma * sigma);
	else if (s.window == HCS_WINDOW_TUKEY)
		rtukey = 1.f / sigma * S * S;
	s.weights.assign((size_t)S * S, 1.0f);
	for (int i = 0; i < S; ++i)
		for (int j = 0; j < S; ++j) {
			float weight = 1.0f;
			if (s.window == HCS_WINDOW_GAUSS) {
				float dist = std::hypot(di_factor * i + sub_halfwidth, di_factor * j + sub_halfwidth) / max_dist;
				weight     = std::exp((double)(-(dist * dist) * rsigma_squared));
			} else if (s.window == HCS_WINDOW_TUKEY) {
				if (S / 2 - std::abs(S / 2 - i) <= sigma * S / 2)
					weight *= 0.5f * (1.f - cosf(2.f * M_PI * (S / 2 - std::abs(S / 2 - i)) * rtukey));
				if (S / 2 - std::abs(S / 2 - j) <= sigma * S / 2)
					weight *= 0.5f * (1.f - cosf(2.f * M_PI * (S / 2 - std::abs(S / 2 - j)) * rtukey));
			} else if (s.window == HCS_WINDOW_SQUARE) {
				float inv_dist = 1 - std::hypot(di_factor * i + sub_halfwidth, di_factor * j + sub_halfwidth) / max_dist;
				weight         = inv_dist * inv_dist;
			}
			s.weights[(size_t)i * S + j] = weight;
		}
	SensorDev d{};
	d.geom = s.geom, d.cx = s.cx, d.cy = s.cy, d.S = S, d.window = s.window, d.sigma = sigma;
	d.resolution = resolution;
	for (int a = 0; a < 3; ++a)
		d.size[a] = g.size[a];
	d.rmean = rmean, d.rS = rS;
	float *w = dalloc<float>(c->step_allocs, s.weights.size());
	CK(cudaMemcpyAsync(w, s.weights.data(), s.weights.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
	d.weights    = w;
	size_t ncell = (size_t)c->cfg.n_envs * s.cx * s.cy;
	d.image      = dalloc<float>(c->step_allocs, ncell);
	CK(cudaMemsetAsync(d.image, 0, ncell * sizeof(float), c->stream));
	d.bin_count  = dalloc<int32_t>(c->step_allocs, ncell);
	d.bin_offset = dalloc<int32_t>(c->step_allocs, ncell + 1);
	d.bin_cursor = dalloc<int32_t>(c->step_allocs, ncell);
	d.scan_tmp   = dalloc<int32_t>(c->step_allocs, ncell / 1024 + 2);
	d.raster_counter = dalloc<int32_t>(c->step_allocs, 1);
	s.dev        = d; // items are sized in finalize() once the triangle pool capacity is known
	CK(cudaMallocHost((void **)&s.h_image, std::max<size_t>(ncell, 1) * sizeof(float)));
}

template <class T>
static T *upload(hcs_ctx *c, const std::vector<T> &v)
{
	T *p = dalloc<T>(c->step_allocs, v.size());
	if (!v.empty())
		CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
	return p;
}

// Static ray grid of a curved sensor (geom frame) + its per-step buffers.  Only cells that hold rays get an id;
// the per-(env, cell) triangle bins are sized like the flat sensor's.
static void build_curved(hcs_ctx *c, CurvedHost &s, int max_tris)
{
	const int n_env = c->cfg.n_envs, nr = s.n_rays(), nt = s.n_taxels();
	CurvedDev d{};
	d.geom = s.geom, d.n_rays = nr, d.n_taxels = nt, d.include_margin = s.include_margin;
	double lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
	for (int r = 0; r < nr; ++r)
		for (int a = 0; a < 3; ++a) {
			double v = s.ray_pos[3 * (size_t)r + a];
			lo[a]    = r == 0 ? v : std::min(lo[a], v);
			hi[a]    = r == 0 ? v : std::max(hi[a], v);
		}
	double ext = std::max({ hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2] });
	d.cell     = std::max({ 2 * s.include_margin, ext / 48, 1e-9 });
	for (int a = 0; a < 3; ++a) {
		d.origin[a] = lo[a] - 1e-9;
		d.dims[a]   = (int)std::floor((hi[a] - d.origin[a]) / d.cell) + 1;
	}
	std::vector<int32_t> lookup((size_t)d.dims[0] * d.dims[1] * d.dims[2], -1), ray_cell(nr);
	int n_cells = 0;
	for (int r = 0; r < nr; ++r) {
		int ci[3];
		for (int a = 0; a < 3; ++a)
			ci[a] = std::min(d.dims[a] - 1, std::max(0, (int)std::floor((s.ray_pos[3 * (size_t)r + a] - d.origin[a]) / d.cell)));
		size_t idx = ((size_t)ci[0] * d.dims[1] + ci[1]) * d.dims[2] + ci[2];
		if (lookup[idx] < 0)
			lookup[idx] = n_cells++;
		ray_cell[r] = lookup[idx];
	}
	d.n_cells = n_cells;
	std::vector<int32_t> cell_off(n_cells + 1, 0), cell_rays(nr);
	for (int r = 0; r < nr; ++r)
		cell_off[ray_cell[r] + 1]++;
	for (int k = 0; k < n_cells; ++k)
		cell_off[k + 1] += cell_off[k];
	{
		std::vector<int32_t> cur(cell_off.begin(), cell_off.end() - 1);
		for (int r = 0; r < nr; ++r)
			cell_rays[cur[ray_cell[r]]++] = r;
	}
	d.ray_pos      = upload(c, s.ray_pos);
	d.ray_nrm      = upload(c, s.ray_nrm);
	d.cell_lookup  = upload(c, lookup);
	d.cell_ray_off = upload(c, cell_off);
	d.cell_rays    = upload(c, cell_rays);
	d.taxel_off    = upload(c, s.taxel_off);
	d.taxel_ray    = upload(c, s.taxel_ray);
	d.taxel_w      = upload(c, s.taxel_w);
	size_t ncell   = (size_t)n_env * n_cells;
	d.bin_count    = dalloc<int32_t>(c->step_allocs, ncell);
	d.bin_offset   = dalloc<int32_t>(c->step_allocs, ncell + 1);
	d.bin_cursor   = dalloc<int32_t>(c->step_allocs, ncell);
	d.scan_tmp     = dalloc<int32_t>(c->step_allocs, ncell / 1024 + 2);
	size_t cap     = std::min<size_t>(std::max<size_t>(8 * (size_t)max_tris, 32 * ncell), (size_t)1 << 30);
	d.items_cap    = (int)cap;
	d.bin_items    = dalloc<int32_t>(c->step_allocs, cap);
	d.raw          = dalloc<double>(c->step_allocs, (size_t)n_env * nr);
	d.values       = dalloc<float>(c->step_allocs, (size_t)n_env * nt);
	CK(cudaMemsetAsync(d.raw, 0, std::max<size_t>((size_t)n_env * nr, 1) * sizeof(double), c->stream));
	CK(cudaMemsetAsync(d.values, 0, std::max<size_t>((size_t)n_env * nt, 1) * sizeof(float), c->stream));
	CK(cudaMallocHost((void **)&s.h_values, std::max<size_t>((size_t)n_env * nt, 1) * sizeof(float)));
	memset(s.h_values, 0, std::max<size_t>((size_t)n_env * nt, 1) * sizeof(float));
	s.dev = d;
}

// hcs_step replays a CUDA graph when the step's inputs are at most this large (they are staged through a pinned buffer of
// the context, so that the graph's copy node has fixed addresses): the reference's own case, one mjData per step, where
// launch overheads are most of the call
constexpr size_t GRAPH_INPUT_BYTES = 256u << 10;

static void release_step_buffers(hcs_ctx *c)
{
	free_bag(c->step_allocs);
	for (SensorHost &s : c->sensors)
		if (s.h_image) {
			cudaFreeHost(s.h_image);
			s.h_image = nullptr;
		}
	for (CurvedHost &s : c->curved)
		if (s.h_values) {
			cudaFreeHost(s.h_values);
			s.h_values = nullptr;
		}
	for (TaxelHost &s : c->taxel)
		if (s.h_values) {
			cudaFreeHost(s.h_values);
			s.h_values = nullptr;
		}
	if (c->h_pair)
		cudaFreeHost(c->h_pair), c->h_pair = nullptr;
	if (c->h_wrench)
		cudaFreeHost(c->h_wrench), c->h_wrench = nullptr;
	if (c->h_flags)
		cudaFreeHost(c->h_flags), c->h_flags = nullptr;
	if (c->h_in)
		cudaFreeHost(c->h_in), c->h_in = nullptr;
	for (auto &row : c->step_graph)
		for (cudaGraphExec_t &g : row)
			if (g)
				cudaGraphExecDestroy(g), g = nullptr;
	c->steps_since_finalize = 0;
	for (hcs_ctx::Slot &sl : c->slot)
		if (sl.h_flags)
			cudaFreeHost(sl.h_flags), sl.h_flags = nullptr, sl.dh_flags = nullptr;
	c->d_pairs = nullptr;
}

static void finalize(hcs_ctx *c)
{
	// not finalized until the very end: if anything below throws (mesh rebuild, cudaMalloc, a sensor limit) the step
	// entry points refuse to run on the freed buffers instead of launching kernels on them
	c->finalized = false;
	release_step_buffers(c);
	const int n_env = c->cfg.n_envs, ng = (int)c->geoms.size(), np = (int)c->pairs.size();
	for (GeomHost &g : c->geoms)
		if (!g.uploaded) {
			upload_geom(c, g);
			g.uploaded = true;
		}
	for (SensorHost &s : c->sensors) {
		if (!c->geoms[s.geom].env_sizes.empty())
			throw std::runtime_error("per-environment sizes: a flat-sensor geom keeps one size (its taxel grid is laid out from it)");
		build_sensor(c, s);
	}
	c->n_counters = 6 + PAIR_COUNTERS * (size_t)np;
	c->d_counters = dalloc<int32_t>(c->step_allocs, 2 * c->n_counters);
	CK(cudaMemsetAsync(c->d_counters, 0, 2 * c->n_counters * sizeof(int32_t), c->stream));
	c->set_clean[0] = c->set_clean[1] = true;
	c->cur_set                        = 0;
	build_pairs(c);
	StepIO io{};
	io.n_env = n_env, io.n_geoms = ng, io.n_pairs = np;
	CK(cudaDeviceGetAttribute(&io.n_sms, cudaDevAttrMultiProcessorCount, c->cfg.device));
	io.representation = c->cfg.representation;
	io.apply_forces   = c->cfg.apply_contact_forces;
	io.flags          = c->d_counters;
	io.max_faces      = std::max(0, c->cfg.max_faces);
	io.faces          = dalloc<hcs_face>(c->step_allocs, io.max_faces);
	io.face_count     = c->d_counters + 4;
	io.face_verts     = c->cfg.face_vertices && io.max_faces > 0 ?
	                        dalloc<double>(c->step_allocs, (size_t)io.max_faces * HCS_FACE_VERTEX_STRIDE) : nullptr;
	bool tactile      = false;
	for (const PairDesc &P : c->pair_desc)
		tactile |= P.emit_tactile != 0;
	io.max_tris = 0;
	if (tactile) {
		long automatic = 0;
		for (const PairDesc &P : c->pair_desc)
			if (P.emit_tactile) // every tet could emit one polygon of <= 8 fan triangles per env; cap the default
				automatic += std::min<long>(8L * P.nq, 4096);
		io.max_tris = c->cfg.max_tactile_triangles > 0 ? c->cfg.max_tactile_triangles :
		                                                   (int)std::min<long>(automatic * n_env, 1L << 28);
	}
	for (SensorHost &sh : c->sensors) { // bins hold (triangle, taxel) overlaps back to back
		size_t ncell = (size_t)n_env * sh.cx * sh.cy;
		size_t per   = c->cfg.max_triangles_per_taxel > 0 ? c->cfg.max_triangles_per_taxel : 32;
		size_t cap   = std::min<size_t>(std::max<size_t>(16 * (size_t)io.max_tris, per * ncell), (size_t)1 << 30);
		sh.dev.items_cap = (int)cap;
		sh.dev.bin_items = dalloc<int32_t>(c->step_allocs, cap);
	}
	for (CurvedHost &ch : c->curved)
		build_curved(c, ch, io.max_tris);
	io.tri_vd = nullptr;
	if (!c->taxel.empty() && io.max_tris > 0)
		io.tri_vd = dalloc<double>(c->step_allocs, (size_t)9 * io.max_tris);
	io.tri_elem = nullptr;
	for (const TaxelHost &th : c->taxel)
		if (th.sample_method == 1 && io.max_tris > 0 && !io.tri_elem)
			io.tri_elem = dalloc<uint2>(c->step_allocs, (size_t)io.max_tris);
	for (TaxelHost &th : c->taxel) {
		const int nt = th.n_taxels();
		TaxelDev d{};
		d.geom = th.geom, d.n_taxels = nt, d.method = th.method, d.visualize = th.visualize;
		d.sample_method = th.sample_method;
		if (th.sample_method == 1 && c->pairs.size() > 512)
			throw std::runtime_error("taxel sensor with sample_method area_importance: at most 512 geom pairs (sort key layout)");
		if (th.sample_method == 1) { // one stratum of sample_resolution * total_area per sample and surface
			long per_surface = (long)std::ceil(1.0 / th.sample_resolution) + 2;
			d.max_samples    = (int)std::min<long>(std::max<long>(per_surface * (long)std::max<size_t>(c->pairs.size(), 1), 16), 1L << 20);
			d.samples        = dalloc<double>(c->step_allocs, (size_t)n_env * d.max_samples * 4);
			d.n_samples      = dalloc<int32_t>(c->step_allocs, n_env);
			d.env_offset     = dalloc<int32_t>(c->step_allocs, (size_t)n_env + 1);
			d.env_cursor     = dalloc<int32_t>(c->step_allocs, n_env);
			d.env_items      = dalloc<int32_t>(c->step_allocs, (size_t)std::max(io.max_tris, 1));
		}
		d.include_margin = th.include_margin, d.sample_resolution = th.sample_resolution;
		d.taxel_pos   = upload(c, th.taxel_pos);
		size_t ncell  = (size_t)n_env * nt;
		d.values      = dalloc<float>(c->step_allocs, ncell);
		d.env_tris    = dalloc<int32_t>(c->step_allocs, n_env);
		d.bin_count   = dalloc<int32_t>(c->step_allocs, ncell);
		d.bin_offset  = dalloc<int32_t>(c->step_allocs, ncell + 1);
		d.bin_cursor  = dalloc<int32_t>(c->step_allocs, ncell);
		d.scan_tmp    = dalloc<int32_t>(c->step_allocs, ncell / 1024 + 2);
		size_t cap    = std::min<size_t>(std::max<size_t>(16 * (size_t)io.max_tris, 32 * ncell), (size_t)1 << 30);
		d.items_cap   = (int)cap;
		d.bin_items   = dalloc<int32_t>(c->step_allocs, cap);
		CK(cudaMemsetAsync(d.values, 0, std::max<size_t>(ncell, 1) * sizeof(float), c->stream)); // channel.values.resize(n)
		CK(cudaMallocHost((void **)&th.h_values, std::max<size_t>(ncell, 1) * sizeof(float)));
		memset(th.h_values, 0, std::max<size_t>(ncell, 1) * sizeof(float));
		th.dev = d;
	}
	c->sensor_dev.clear();
	for (SensorHost &sh : c->sensors)
		c->sensor_dev.push_back(sh.dev);
	c->d_sensors = dalloc<SensorDev>(c->step_allocs, c->sensor_dev.size());
	if (!c->sensor_dev.empty())
		CK(cudaMemcpyAsync(c->d_sensors, c->sensor_dev.data(), c->sensor_dev.size() * sizeof(SensorDev),
		                   cudaMemcpyHostToDevice, c->stream));
	io.tri_pool    = dalloc<TactileTri>(c->step_allocs, io.max_tris);
	io.tri_count   = c->d_counters + 5;
	io.pair_out    = dalloc<hcs_pair_result>(c->step_allocs, (size_t)n_env * np);
	io.geom_wrench = dalloc<double>(c->step_allocs, (size_t)n_env * ng * 6);
	CK(cudaMemsetAsync(io.pair_out, 0, (size_t)n_env * np * sizeof(hcs_pair_result), c->stream));
	CK(cudaMemsetAsync(io.geom_wrench, 0, (size_t)n_env * ng * 6 * sizeof(double), c->stream));
	c->d_xpos = dalloc<double>(c->step_allocs, (size_t)n_env * ng * 18); // xpos | xmat | vel back to back: one copy suffices
	c->d_xmat = c->d_xpos + (size_t)n_env * ng * 3;
	c->d_vel  = c->d_xmat + (size_t)n_env * ng * 9;
	if ((size_t)n_env * ng * 18 * sizeof(double) <= GRAPH_INPUT_BYTES)
		CK(cudaMallocHost((void **)&c->h_in, std::max<size_t>((size_t)n_env * ng * 18, 1) * sizeof(double)));
	CK(cudaMallocHost((void **)&c->h_pair, std::max<size_t>((size_t)n_env * np, 1) * sizeof(hcs_pair_result)));
	CK(cudaMallocHost((void **)&c->h_wrench, std::max<size_t>((size_t)n_env * ng * 6, 1) * sizeof(double)));
	CK(cudaMallocHost((void **)&c->h_flags, 4 * sizeof(int32_t)));
	memset(c->h_flags, 0, 4 * sizeof(int32_t));
	for (hcs_ctx::Slot &sl : c->slot) {
		sl.d_xpos   = dalloc<double>(c->step_allocs, (size_t)n_env * ng * 3);
		sl.d_xmat   = dalloc<double>(c->step_allocs, (size_t)n_env * ng * 9);
		sl.d_vel    = dalloc<double>(c->step_allocs, (size_t)n_env * ng * 6);
		sl.d_wrench = dalloc<double>(c->step_allocs, (size_t)n_env * ng * 6);
		CK(cudaMallocHost((void **)&sl.h_flags, 4 * sizeof(int32_t)));
		memset(sl.h_flags, 0, 4 * sizeof(int32_t));
		if (cudaHostGetDevicePointer((void **)&sl.dh_flags, sl.h_flags, 0) != cudaSuccess) {
			sl.dh_flags = nullptr;
			cudaGetLastError();
		}
		sl.ticket = -1, sl.waited = true, sl.status = HCS_OK;
	}
	// device addresses of the pinned mirrors (pinned allocations are mapped under unified addressing)
	c->dh_wrench = nullptr, c->dh_flags = nullptr;
	if (cudaHostGetDevicePointer((void **)&c->dh_wrench, c->h_wrench, 0) != cudaSuccess ||
	    cudaHostGetDevicePointer((void **)&c->dh_flags, c->h_flags, 0) != cudaSuccess) {
		c->dh_wrench = nullptr, c->dh_flags = nullptr;
		cudaGetLastError();
	}
	io.geom_wrench_host = nullptr, io.flags_host = nullptr;
	c->io = io;
	CK(cudaStreamSynchronize(c->stream));
	c->finalized = true;
}

// Point the context's records (c->io, c->pair_desc: what the launches and the getters use) at one set of step counters
static void select_counter_set(hcs_ctx *c, int set)
{
	int32_t *base   = c->d_counters + (size_t)set * c->n_counters;
	c->io.flags      = base;
	c->io.face_count = base + 4;
	c->io.tri_count  = base + 5;
	for (size_t pi = 0; pi < c->pair_desc.size(); ++pi)
		if (c->pair_desc[pi].kind != PAIR_NONE)
			c->pair_desc[pi].counters = base + 6 + PAIR_COUNTERS * pi;
	c->cur_set = set;
}

// the counter set the next step will use: a clean one, else -1 (set 0 behind a memset)
static int next_counter_set(const hcs_ctx *c)
{
	static const bool no_memset_ok = getenv("HCS_STEP_MEMSET") == nullptr; // HCS_STEP_MEMSET=1: a memset per step (A/B)
	if (!no_memset_ok)
		return -1;
	return c->set_clean[0] ? 0 : (c->set_clean[1] ? 1 : -1);
}

// HCS_DEBUG_LAUNCH=1: name the launch of a step that the runtime refused (launches are not checked one by one otherwise)
static void dbg_launch(const char *what, int pair)
{
	static const bool on = getenv("HCS_DEBUG_LAUNCH") != nullptr;
	if (!on)
		return;
	const cudaError_t e = cudaPeekAtLastError();
	int dev = -1;
	cudaGetDevice(&dev);
	if (e != cudaSuccess)
		fprintf(stderr, "hcs: %s (pair %d, current device %d): %s\n", what, pair, dev, cudaGetErrorString(e));
}

// Every step zeroes the other set behind itself, so consecutive steps alternate between the sets and none needs a memset
// (a captured CUDA graph bakes its set in: hcs_step keeps one graph per set).
static void step_device(hcs_ctx *c, const double *xpos, const double *xmat, const double *vel, int with_sensors,
                        bool direct_out = false, double *wrench_dev = nullptr, int32_t *flags_mapped = nullptr)
{
	int set = next_counter_set(c);
	const bool memset_first = set < 0;
	if (memset_first)
		set = 0;
	select_counter_set(c, set);
	StepIO io = c->io;
	io.zero_next = c->d_counters + (size_t)(1 - set) * c->n_counters;
	io.n_zero    = (int)c->n_counters;
	io.xpos = xpos, io.xmat = xmat, io.vel = vel;
	// direct_out: the finalize kernel also writes wrenches and flags into the context's mapped pinned mirrors
	io.geom_wrench_host = direct_out ? c->dh_wrench : nullptr;
	io.flags_host       = direct_out ? c->dh_flags : nullptr;
	if (wrench_dev) // pipelined steps: per-slot device buffer for the wrenches, per-slot mapped flags
		io.geom_wrench = wrench_dev;
	if (flags_mapped)
		io.flags_host = flags_mapped;
	cudaStream_t s = c->stream;
	int64_t k      = 0;
	bool prof      = c->profiling;
	{ // an error some earlier, unchecked runtime call of this host thread left behind must not be blamed on this step's launches
		const cudaError_t stale = cudaGetLastError();
		if (stale != cudaSuccess && getenv("HCS_DEBUG_LAUNCH"))
			fprintf(stderr, "hcs: stale CUDA error before the step: %s\n", cudaGetErrorString(stale));
	}
	if (prof)
		CK(cudaEventRecord(c->ev[0], s));
	if (memset_first) // flags, pool counts, flat-list counters
		CK(cudaMemsetAsync(c->d_counters, 0, c->n_counters * sizeof(int32_t), s));
	c->set_clean[set]     = false;
	c->set_clean[1 - set] = true; // (by this step's finalize kernel)
	if (prof)
		CK(cudaEventRecord(c->ev[1], s));
	int n_active = 0;
	for (const PairDesc &P : c->pair_desc)
		n_active += P.kind != PAIR_NONE;
	// Scenes with several pairs: each pair's broadphase -> narrowphase chain goes to one of N_AUX side streams, so
	// the pairs' kernels overlap (launch latencies, ramp-up and tails of one pair are filled by the others) instead
	// of running as 2 x n_pairs serialised launches; the main stream joins them before the finalize.  With stage
	// profiling on, everything stays on the main stream so that the stage events mean what they say.
	static const bool allow_fork = getenv("HCS_NO_FORK") == nullptr;
	const bool forked = allow_fork && !prof && n_active > 1;
	if (forked) {
		CK(cudaEventRecord(c->ev_fork, s));
		int j = 0;
		for (const PairDesc &P : c->pair_desc)
			if (P.kind != PAIR_NONE) {
				cudaStream_t st = c->aux[j % hcs_ctx::N_AUX];
				if (j < hcs_ctx::N_AUX)
					CK(cudaStreamWaitEvent(st, c->ev_fork, 0));
				launch_broadphase(P, io, st);
				dbg_launch("forked broadphase", P.index);
				launch_narrowphase(P, io, st, /*chained=*/true);
				dbg_launch("forked narrowphase", P.index);
				k += 2;
				++j;
			}
		for (int i = 0; i < std::min(j, (int)hcs_ctx::N_AUX); ++i) {
			CK(cudaEventRecord(c->ev_join[i], c->aux[i]));
			CK(cudaStreamWaitEvent(s, c->ev_join[i], 0));
		}
	} else {
		for (const PairDesc &P : c->pair_desc)
			if (P.kind != PAIR_NONE) {
				launch_broadphase(P, io, s);
				dbg_launch("broadphase", P.index);
				++k;
			}
		if (prof)
			CK(cudaEventRecord(c->ev[2], s));
		for (const PairDesc &P : c->pair_desc)
			if (P.kind != PAIR_NONE) {
				launch_narrowphase(P, io, s, /*chained=*/!prof);
				dbg_launch("narrowphase", P.index);
				++k;
			}
		if (prof)
			CK(cudaEventRecord(c->ev[3], s));
	}
	dbg_launch("before finalize", -1);
	k += launch_finalize(c->d_pairs, io, s, /*chained=*/!prof && !forked);
	dbg_launch("finalize", -1);
	if (prof)
		CK(cudaEventRecord(c->ev[4], s));
	if (with_sensors) {
		k += launch_tactile(c->sensor_dev.data(), c->d_sensors, (int)c->sensor_dev.size(), io, c->d_pairs, s);
		for (CurvedHost &ch : c->curved)
			k += launch_curved(ch.dev, io, c->d_pairs, s);
		for (TaxelHost &th : c->taxel)
			k += launch_taxel(th.dev, io, c->d_pairs, s);
	}
	if (prof)
		CK(cudaEventRecord(c->ev[5], s));
	CK(cudaGetLastError());
	c->kernels_last_step = k;
	c->step_counter++;
	c->results_on_host = c->pairs_on_host = c->sensors_on_host = false;
	c->last_with_sensors                   = with_sensors != 0;
}

// D2H of what the caller applies: per-geom wrenches, flags and (when computed) the sensor outputs.  The per-pair
// results are diagnostics (the reference has no such output): they are copied only when asked for
// (hcs_get_pair_results, hcs_get_counters, hcs_fetch_results), so hcs_step does not pay for them every step.
static void fetch_enqueue(hcs_ctx *c, int with_sensors, bool with_pairs)
{
	const int n_env = c->cfg.n_envs, ng = (int)c->geoms.size(), np = (int)c->pairs.size();
	cudaStream_t s = c->stream;
	if (with_pairs)
		CK(cudaMemcpyAsync(c->h_pair, c->io.pair_out, (size_t)n_env * np * sizeof(hcs_pair_result), cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(c->h_wrench, c->io.geom_wrench, (size_t)n_env * ng * 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(c->h_flags, c->io.flags, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
	if (with_sensors)
		for (SensorHost &sh : c->sensors)
			CK(cudaMemcpyAsync(sh.h_image, sh.dev.image, (size_t)n_env * sh.cx * sh.cy * sizeof(float), cudaMemcpyDeviceToHost, s));
	if (with_sensors)
		for (CurvedHost &ch : c->curved)
			CK(cudaMemcpyAsync(ch.h_values, ch.dev.values, (size_t)n_env * ch.n_taxels() * sizeof(float),
			                   cudaMemcpyDeviceToHost, s));
	if (with_sensors)
		for (TaxelHost &th : c->taxel)
			CK(cudaMemcpyAsync(th.h_values, th.dev.values, (size_t)n_env * th.n_taxels() * sizeof(float),
			                   cudaMemcpyDeviceToHost, s));
}

static void fetch(hcs_ctx *c, int with_sensors, bool with_pairs = true)
{
	fetch_enqueue(c, with_sensors, with_pairs);
	CK(cudaStreamSynchronize(c->stream));
	c->results_on_host = true;
	c->pairs_on_host   = c->pairs_on_host || with_pairs;
	c->sensors_on_host = with_sensors != 0;
}

static int check_flags(hcs_ctx *c)
{
	if (c->h_flags[1]) {
		c->err = "broadphase work queue overflow (shared-memory LIFO queues of a unit full)";
		return HCS_E_CAPACITY;
	}
	if (c->h_flags[0] & 1) {
		c->err = "broadphase candidate overflow (unused since the per-unit slabs were removed)";
		return HCS_E_CAPACITY;
	}
	if (c->h_flags[0] & 2) {
		c->err = "tactile triangle pool overflow: raise hcs_config.max_tactile_triangles";
		return HCS_E_CAPACITY;
	}
	if (c->h_flags[0] & 4) {
		c->err = "tactile bin overflow: raise hcs_config.max_triangles_per_taxel (average bin depth); for a taxel sensor with "
		         "sample_method area_importance: more than 4096 contact-surface triangles in one environment, or more samples "
		         "than 1 / sample_resolution + 2 per surface";
		return HCS_E_CAPACITY;
	}
	if (c->h_flags[0] & 8) {
		c->err = "candidate pool overflow: raise hcs_config.max_candidates_per_slice (average candidates per (env, slice) unit) "
		         "or HCS_MAX_TOTAL_CANDIDATES (entries per pair)";
		return HCS_E_CAPACITY;
	}
	if (c->h_flags[0] & 16) {
		c->err = "a contact polygon's contribution (force, torque, area) was not finite or beyond 2^26 SI units: the exact "
		         "accumulators cannot hold it (check poses and velocities for NaN / infinity)";
		return HCS_E_INVALID;
	}
	return HCS_OK;
}

} // namespace hcs

// =====================================================================================================
// C ABI
// =====================================================================================================
// Every entry point that touches CUDA runs with the context's device current and restores the caller's device on
// the way out: contexts on different GPUs can be driven from one thread (or from one thread per device) in any order.
struct DeviceGuard {
	int prev = -1, dev;
	explicit DeviceGuard(int d) : dev(d)
	{
		if (cudaGetDevice(&prev) != cudaSuccess)
			prev = -1;
		if (prev != dev && cudaSetDevice(dev) != cudaSuccess)
			throw std::runtime_error("cudaSetDevice failed for the context's device");
	}
	~DeviceGuard()
	{
		if (prev >= 0 && prev != dev)
			cudaSetDevice(prev);
	}
};

#define API_BEGIN(ctx_)            \
	if (!(ctx_))                   \
		return HCS_E_INVALID;      \
	try {                          \
		DeviceGuard device_guard_((ctx_)->cfg.device);
#define API_END(ctx_)                        \
	}                                        \
	catch (const std::exception &e)          \
	{                                        \
		(ctx_)->err = e.what();              \
		return HCS_E_CUDA;                   \
	}

extern "C" {

const char *hcs_version(void) { return "hcs_b200 0.1 (sm_100a, fp64 geometry mode)"; }

const char *hcs_last_error(const hcs_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hcs_create(const hcs_config *cfg, hcs_ctx **out)
{
	if (!cfg || !out || cfg->n_envs < 1) {
		g_create_error = "hcs_create: bad arguments";
		return HCS_E_INVALID;
	}
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		// no CPU fallback: the engine exists only on the GPU
		g_create_error = std::string("hcs_create: no usable CUDA device (") + cudaGetErrorString(e) + ")";
		return HCS_E_CUDA;
	}
	if (cfg->device < 0 || cfg->device >= ndev) {
		g_create_error = "hcs_create: device ordinal out of range";
		return HCS_E_INVALID;
	}
	hcs_ctx *c = new hcs_ctx();
	c->cfg     = *cfg;
	try {
		CK(cudaSetDevice(cfg->device));
		if (cfg->stream) {
			c->stream = (cudaStream_t)cfg->stream;
		} else {
			CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
			c->own_stream = true;
		}
		for (auto &ev : c->ev)
			CK(cudaEventCreate(&ev));
		for (int i = 0; i < hcs_ctx::N_AUX; ++i) { // side streams for the per-pair kernels of multi-pair scenes
			CK(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
		}
		CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
		CK(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
		for (hcs_ctx::Slot &sl : c->slot)
			for (cudaEvent_t *e : { &sl.ev_in, &sl.ev_free, &sl.ev_kdone, &sl.ev_done })
				CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
	} catch (const std::exception &ex) {
		g_create_error = ex.what();
		delete c;
		return HCS_E_CUDA;
	}
	*out = c;
	return HCS_OK;
}

void hcs_destroy(hcs_ctx *c)
{
	if (!c)
		return;
	cudaSetDevice(c->cfg.device);
	cudaStreamSynchronize(c->stream);
	release_step_buffers(c);
	for (GeomHost &g : c->geoms)
		free_bag(g.allocs);
	for (auto &ev : c->ev)
		if (ev)
			cudaEventDestroy(ev);
	for (int i = 0; i < hcs_ctx::N_AUX; ++i) {
		if (c->aux[i])
			cudaStreamDestroy(c->aux[i]);
		if (c->ev_join[i])
			cudaEventDestroy(c->ev_join[i]);
	}
	if (c->ev_fork)
		cudaEventDestroy(c->ev_fork);
	for (hcs_ctx::Slot &sl : c->slot)
		for (cudaEvent_t e : { sl.ev_in, sl.ev_free, sl.ev_kdone, sl.ev_done })
			if (e)
				cudaEventDestroy(e);
	if (c->copy_in)
		cudaStreamDestroy(c->copy_in);
	if (c->copy_out)
		cudaStreamDestroy(c->copy_out);
	if (c->own_stream)
		cudaStreamDestroy(c->stream);
	delete c;
}

int hcs_add_geom(hcs_ctx *c, int mj_geom_type, const double size[3], const float *mesh_vert, int n_vert,
                 const int32_t *mesh_face, int n_face, const double props[5])
{
	API_BEGIN(c)
	if (!size || !props) {
		c->err = "hcs_add_geom: size/props required";
		return HCS_E_INVALID;
	}
	if (mj_geom_type == HCS_GEOM_MESH && mesh_face)
		for (size_t i = 0; i < 3 * (size_t)std::max(n_face, 0); ++i)
			if (mesh_face[i] < 0 || mesh_face[i] >= n_vert) {
				c->err = "hcs_add_geom: mesh face index out of range";
				return HCS_E_INVALID;
			}
	GeomHost g;
	g.mj_type = mj_geom_type;
	for (int i = 0; i < 3; ++i)
		g.size[i] = size[i];
	for (int i = 0; i < 5; ++i)
		g.props[i] = props[i];
	std::string err;
	if (!build_geom_mesh(mj_geom_type, size, mesh_vert, n_vert, mesh_face, n_face, props, g.mesh, err)) {
		c->err = err;
		return HCS_E_UNSUPPORTED;
	}
	c->geoms.push_back(std::move(g));
	c->finalized = false;
	return (int)c->geoms.size() - 1;
	API_END(c)
}

int hcs_add_soft_mesh(hcs_ctx *c, const double *verts, int n_vert, const int32_t *tets, int n_tet,
                      const double *vertex_pressure, const double props[5])
{
	API_BEGIN(c)
	if (!verts || !tets || !vertex_pressure || !props || n_vert < 4 || n_tet < 1 || !(props[0] > 0)) {
		c->err = "hcs_add_soft_mesh: bad arguments (modulus must be > 0)";
		return HCS_E_INVALID;
	}
	for (size_t i = 0; i < 4 * (size_t)n_tet; ++i)
		if (tets[i] < 0 || tets[i] >= n_vert) {
			c->err = "hcs_add_soft_mesh: tet vertex index out of range";
			return HCS_E_INVALID;
		}
	GeomHost g;
	g.mj_type = HCS_GEOM_MESH;
	g.custom  = true;
	for (int i = 0; i < 5; ++i)
		g.props[i] = props[i];
	g.mesh.soft = true;
	g.mesh.verts.assign(verts, verts + 3 * (size_t)n_vert);
	g.mesh.elems.assign(tets, tets + 4 * (size_t)n_tet);
	g.mesh.pressure.assign(vertex_pressure, vertex_pressure + n_vert);
	c->geoms.push_back(std::move(g));
	c->finalized = false;
	return (int)c->geoms.size() - 1;
	API_END(c)
}

int hcs_add_rigid_mesh(hcs_ctx *c, const double *verts, int n_vert, const int32_t *tris, int n_tri, const double props[5])
{
	API_BEGIN(c)
	if (!verts || !tris || !props || n_vert < 3 || n_tri < 1) {
		c->err = "hcs_add_rigid_mesh: bad arguments";
		return HCS_E_INVALID;
	}
	for (size_t i = 0; i < 3 * (size_t)n_tri; ++i)
		if (tris[i] < 0 || tris[i] >= n_vert) {
			c->err = "hcs_add_rigid_mesh: triangle vertex index out of range";
			return HCS_E_INVALID;
		}
	GeomHost g;
	g.mj_type = HCS_GEOM_MESH;
	g.custom  = true;
	for (int i = 0; i < 5; ++i)
		g.props[i] = props[i];
	g.props[0]  = 0;
	g.mesh.soft = false;
	g.mesh.verts.assign(verts, verts + 3 * (size_t)n_vert);
	g.mesh.elems.assign(tris, tris + 3 * (size_t)n_tri);
	c->geoms.push_back(std::move(g));
	c->finalized = false;
	return (int)c->geoms.size() - 1;
	API_END(c)
}

int hcs_update_geom(hcs_ctx *c, int geom, const double size[3])
{
	API_BEGIN(c)
	if (geom < 0 || geom >= (int)c->geoms.size() || !size || c->geoms[geom].custom) {
		c->err = "hcs_update_geom: bad geom";
		return HCS_E_INVALID;
	}
	GeomHost &g = c->geoms[geom];
	if (g.mj_type == HCS_GEOM_MESH || g.mj_type == HCS_GEOM_PLANE)
		return HCS_OK; // nothing depends on geom_size
	HostMesh m;
	std::string err;
	if (!build_geom_mesh(g.mj_type, size, nullptr, 0, nullptr, 0, g.props, m, err)) {
		c->err = err;
		return HCS_E_UNSUPPORTED;
	}
	for (int i = 0; i < 3; ++i)
		g.size[i] = size[i];
	g.mesh = std::move(m);
	g.uploaded = false;
	g.env_sizes.clear(); // one size for every environment again
	if (c->finalized) { // element counts may change: rebuild the per-pair buffers as well
		CK(cudaStreamSynchronize(c->stream));
		finalize(c);
	}
	return HCS_OK;
	API_END(c)
}

int hcs_set_env_sizes(hcs_ctx *c, int geom, const double *sizes)
{
	API_BEGIN(c)
	if (geom < 0 || geom >= (int)c->geoms.size() || c->geoms[geom].custom) {
		c->err = "hcs_set_env_sizes: bad geom (needs a geom added with hcs_add_geom)";
		return HCS_E_INVALID;
	}
	GeomHost &g = c->geoms[geom];
	if (g.mj_type == HCS_GEOM_MESH || g.mj_type == HCS_GEOM_PLANE) {
		c->err = "hcs_set_env_sizes: mesh and plane geoms do not depend on geom_size";
		return HCS_E_UNSUPPORTED;
	}
	if (!sizes && g.env_sizes.empty())
		return HCS_OK;
	{
		const double *one = sizes ? sizes : g.base_size; // the shared topology is the mesh of environment 0 / the geom's own size
		HostMesh m;
		std::string err;
		if (!build_geom_mesh(g.mj_type, one, nullptr, 0, nullptr, 0, g.props, m, err)) {
			c->err = err;
			return HCS_E_UNSUPPORTED;
		}
		if (sizes && g.env_sizes.empty())
			for (int i = 0; i < 3; ++i)
				g.base_size[i] = g.size[i];
		if (sizes)
			g.env_sizes.assign(sizes, sizes + 3 * (size_t)c->cfg.n_envs);
		else
			g.env_sizes.clear();
		for (int i = 0; i < 3; ++i)
			g.size[i] = one[i];
		g.mesh     = std::move(m);
		g.uploaded = false;
	}
	if (c->finalized) {
		CK(cudaStreamSynchronize(c->stream));
		try {
			finalize(c);
		} catch (const std::exception &ex) {
			c->err = ex.what();
			return std::string(ex.what()).find("per-environment sizes") != std::string::npos ? HCS_E_UNSUPPORTED : HCS_E_CUDA;
		}
	}
	return HCS_OK;
	API_END(c)
}

int hcs_set_pairs(hcs_ctx *c, const int32_t *g1, const int32_t *g2, int n_pairs)
{
	API_BEGIN(c)
	if (n_pairs < 0 || n_pairs > (1 << (32 - TRI_SLICE_BITS)) || (n_pairs > 0 && (!g1 || !g2))) {
		c->err = "hcs_set_pairs: bad arguments";
		return HCS_E_INVALID;
	}
	std::vector<std::pair<int, int>> p;
	for (int i = 0; i < n_pairs; ++i) {
		if (g1[i] < 0 || g2[i] < 0 || g1[i] >= (int)c->geoms.size() || g2[i] >= (int)c->geoms.size() || g1[i] == g2[i]) {
			c->err = "hcs_set_pairs: geom index out of range";
			return HCS_E_INVALID;
		}
		p.emplace_back(g1[i], g2[i]);
	}
	c->pairs     = p;
	c->finalized = false;
	return HCS_OK;
	API_END(c)
}

int hcs_add_flat_sensor(hcs_ctx *c, int geom, double resolution, int sampling_resolution, int window, float sigma)
{
	API_BEGIN(c)
	// the reference reads the three geom_size entries of whatever geom carries the sensor
	// (flat_tactile_sensor.cpp:192-197, 265-267): boxes and ellipsoids have three positive ones
	if (geom < 0 || geom >= (int)c->geoms.size() || !(c->geoms[geom].size[0] > 0) || !(c->geoms[geom].size[1] > 0) ||
	    !(c->geoms[geom].size[2] > 0) || !(resolution > 0) || sampling_resolution < 1 || sampling_resolution > 32 ||
	    window < 0 || window > 3) {
		c->err = "hcs_add_flat_sensor: needs a geom with three positive sizes (box, ellipsoid), resolution > 0, "
		         "1 <= sampling_resolution <= 32, window in 0..3";
		return HCS_E_INVALID;
	}
	SensorHost s{};
	s.geom = geom, s.resolution = resolution, s.S = sampling_resolution, s.window = window, s.sigma = sigma;
	// defaults of flat_tactile_sensor.cpp:148-160
	if (window == HCS_WINDOW_GAUSS && sigma == -1.0f)
		s.sigma = 0.1;
	if (window == HCS_WINDOW_TUKEY && sigma == -1.0f)
		s.sigma = 0.3;
	const double *gs = c->geoms[geom].size;
	s.cx = (int)::floorl(2 * gs[0] / resolution + 0.1); // flat_tactile_sensor.cpp:196-197
	s.cy = (int)::floorl(2 * gs[1] / resolution + 0.1);
	if (s.cx < 1 || s.cy < 1) {
		c->err = "hcs_add_flat_sensor: resolution larger than the sensor geom";
		return HCS_E_INVALID;
	}
	c->sensors.push_back(s);
	c->finalized = false;
	return (int)c->sensors.size() - 1;
	API_END(c)
}

int hcs_update_flat_sensor(hcs_ctx *c, int sensor, int sampling_resolution, int window, float sigma)
{
	API_BEGIN(c)
	if (sensor < 0 || sensor >= (int)c->sensors.size() || sampling_resolution < 1 || sampling_resolution > 32 || window < 0 ||
	    window > 3) {
		c->err = "hcs_update_flat_sensor: needs a flat sensor index, 1 <= sampling_resolution <= 32, window in 0..3";
		return HCS_E_INVALID;
	}
	SensorHost &s = c->sensors[sensor];
	s.S = sampling_resolution, s.window = window, s.sigma = sigma;
	if (window == HCS_WINDOW_GAUSS && sigma == -1.0f) // defaults of flat_tactile_sensor.cpp:148-160
		s.sigma = 0.1;
	if (window == HCS_WINDOW_TUKEY && sigma == -1.0f)
		s.sigma = 0.3;
	if (c->finalized) { // rebuild the step buffers with the new ray grid (rare: a reconfigure request)
		CK(cudaStreamSynchronize(c->stream));
		finalize(c);
	}
	return HCS_OK;
	API_END(c)
}

int hcs_add_curved_sensor(hcs_ctx *c, int geom, int n_taxels, const double *taxel_pos, const double *taxel_nrm,
                          int n_samples, const double *sample_pos, const double *sample_nrm, double include_margin)
{
	API_BEGIN(c)
	if (geom < 0 || geom >= (int)c->geoms.size() || n_taxels < 1 || !taxel_pos || n_samples < 0 ||
	    (n_samples > 0 && (!sample_pos || !sample_nrm)) || !(include_margin > 0)) {
		c->err = "hcs_add_curved_sensor: needs a geom, >= 1 taxel, sample points with normals, include_margin > 0";
		return HCS_E_INVALID;
	}
	CurvedHost s{};
	s.geom = geom, s.include_margin = include_margin;
	s.taxel_pos.assign(taxel_pos, taxel_pos + 3 * (size_t)n_taxels);
	s.taxel_nrm.assign(3 * (size_t)n_taxels, 0.0);
	if (taxel_nrm)
		for (int i = 0; i < n_taxels; ++i) { // normalised like curved_sensor.cpp:187
			const double *t = taxel_nrm + 3 * (size_t)i;
			double l = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
			for (int a = 0; a < 3; ++a)
				s.taxel_nrm[3 * (size_t)i + a] = l > 0 ? t[a] / l : t[a];
		}
	// curved_sensor.cpp:325-368: every sample within include_margin of a taxel (and within 45 degrees of its normal,
	// when it has one) is assigned to it with weight (include_margin - distance)^2; samples are kept in the order
	// they are first used.  dist(i,j) is the reference's (-2 t.s + |t|^2) + |s|^2.
	const double margin_sq = include_margin * include_margin;
	std::vector<std::vector<int32_t>> idx(n_taxels);
	std::vector<std::vector<double>> wgt(n_taxels);
	for (int j = 0; j < n_samples; ++j) {
		const double *sp = sample_pos + 3 * (size_t)j, *sn = sample_nrm + 3 * (size_t)j;
		bool added = false;
		for (int i = 0; i < n_taxels; ++i) {
			const double *t = &s.taxel_pos[3 * (size_t)i], *tn = &s.taxel_nrm[3 * (size_t)i];
			double dist = (-2 * (t[0] * sp[0] + t[1] * sp[1] + t[2] * sp[2]) + (t[0] * t[0] + t[1] * t[1] + t[2] * t[2])) +
			              (sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2]);
			double nsq = tn[0] * tn[0] + tn[1] * tn[1] + tn[2] * tn[2];
			if (dist < margin_sq &&
			    (nsq == 0 || std::acos(tn[0] * sn[0] + tn[1] * sn[1] + tn[2] * sn[2]) < 45 * M_PI / 180.)) {
				if (!added) {
					s.ray_pos.insert(s.ray_pos.end(), sp, sp + 3);
					s.ray_nrm.insert(s.ray_nrm.end(), sn, sn + 3);
					added = true;
				}
				idx[i].push_back(s.n_rays() - 1);
				wgt[i].push_back(std::pow(std::max(0.0, include_margin - std::sqrt(dist)), 2));
			}
		}
	}
	s.taxel_off.assign(1, 0);
	for (int i = 0; i < n_taxels; ++i) {
		s.taxel_ray.insert(s.taxel_ray.end(), idx[i].begin(), idx[i].end());
		s.taxel_w.insert(s.taxel_w.end(), wgt[i].begin(), wgt[i].end());
		s.taxel_off.push_back((int32_t)s.taxel_ray.size());
	}
	c->curved.push_back(std::move(s));
	c->finalized = false;
	return (int)c->curved.size() - 1;
	API_END(c)
}

int hcs_curved_sensor_info(const hcs_ctx *c, int sensor, int *n_taxels, int *n_rays, int *n_assignments)
{
	if (!c || sensor < 0 || sensor >= (int)c->curved.size())
		return HCS_E_INVALID;
	const CurvedHost &s = c->curved[sensor];
	if (n_taxels)
		*n_taxels = s.n_taxels();
	if (n_rays)
		*n_rays = s.n_rays();
	if (n_assignments)
		*n_assignments = (int)s.taxel_ray.size();
	return HCS_OK;
}

int hcs_get_curved_values(hcs_ctx *c, int sensor, float *out)
{
	API_BEGIN(c)
	if (!c->finalized || !out || sensor < 0 || sensor >= (int)c->curved.size())
		return HCS_E_INVALID;
	if (!c->last_with_sensors) {
		c->err = "hcs_get_curved_values: the last step ran with with_sensors == 0";
		return HCS_E_INVALID;
	}
	if (!c->sensors_on_host)
		fetch(c, 1);
	CurvedHost &s = c->curved[sensor];
	memcpy(out, s.h_values, (size_t)c->cfg.n_envs * s.n_taxels() * sizeof(float));
	return HCS_OK;
	API_END(c)
}

const float *hcs_device_curved_values(hcs_ctx *c, int sensor)
{
	if (!c || !c->finalized || sensor < 0 || sensor >= (int)c->curved.size())
		return nullptr;
	return c->curved[sensor].dev.values;
}

int hcs_add_taxel_sensor(hcs_ctx *c, int geom, int n_taxels, const double *taxel_pos, double include_margin,
                         double sample_resolution, int method, int visualize, int sample_method)
{
	API_BEGIN(c)
	if (geom < 0 || geom >= (int)c->geoms.size() || n_taxels < 1 || !taxel_pos || !(include_margin > 0) ||
	    !(sample_resolution > 0) || method < 0 || method > 3 || sample_method < 0 || sample_method > 1) {
		c->err = "hcs_add_taxel_sensor: needs a geom, >= 1 taxel, include_margin > 0, sample_resolution > 0, method in "
		         "0..3, sample_method in 0..1";
		return HCS_E_INVALID;
	}
	TaxelHost s{};
	s.geom = geom, s.method = method, s.visualize = visualize != 0, s.sample_method = sample_method;
	s.include_margin = include_margin, s.sample_resolution = sample_resolution;
	s.taxel_pos.assign(taxel_pos, taxel_pos + 3 * (size_t)n_taxels);
	c->taxel.push_back(std::move(s));
	c->finalized = false;
	return (int)c->taxel.size() - 1;
	API_END(c)
}

int hcs_get_taxel_values(hcs_ctx *c, int sensor, float *out)
{
	API_BEGIN(c)
	if (!c->finalized || !out || sensor < 0 || sensor >= (int)c->taxel.size())
		return HCS_E_INVALID;
	if (!c->last_with_sensors) {
		c->err = "hcs_get_taxel_values: the last step ran with with_sensors == 0";
		return HCS_E_INVALID;
	}
	if (!c->sensors_on_host)
		fetch(c, 1);
	TaxelHost &s = c->taxel[sensor];
	memcpy(out, s.h_values, (size_t)c->cfg.n_envs * s.n_taxels() * sizeof(float));
	return HCS_OK;
	API_END(c)
}

const float *hcs_device_taxel_values(hcs_ctx *c, int sensor)
{
	if (!c || !c->finalized || sensor < 0 || sensor >= (int)c->taxel.size())
		return nullptr;
	return c->taxel[sensor].dev.values;
}

int hcs_sensor_dims(const hcs_ctx *c, int sensor, int *cx, int *cy)
{
	if (!c || sensor < 0 || sensor >= (int)c->sensors.size() || !cx || !cy)
		return HCS_E_INVALID;
	*cx = c->sensors[sensor].cx;
	*cy = c->sensors[sensor].cy;
	return HCS_OK;
}

int hcs_finalize(hcs_ctx *c)
{
	API_BEGIN(c)
	finalize(c);
	return HCS_OK;
	API_END(c)
}

int hcs_step_device(hcs_ctx *c, const double *d_xpos, const double *d_xmat, const double *d_vel, int with_sensors)
{
	API_BEGIN(c)
	if (!c->finalized) {
		c->err = "hcs_step: call hcs_finalize first";
		return HCS_E_NOT_FINALIZED;
	}
	if (!d_xpos || !d_xmat || !d_vel) {
		c->err = "hcs_step_device: null pose/velocity pointer";
		return HCS_E_INVALID;
	}
	step_device(c, d_xpos, d_xmat, d_vel, with_sensors);
	return HCS_OK;
	API_END(c)
}

int hcs_step(hcs_ctx *c, const double *xpos, const double *xmat, const double *vel, int with_sensors)
{
	API_BEGIN(c)
	if (!c->finalized) {
		c->err = "hcs_step: call hcs_finalize first";
		return HCS_E_NOT_FINALIZED;
	}
	if (!xpos || !xmat || !vel) {
		c->err = "hcs_step: null pose/velocity pointer";
		return HCS_E_INVALID;
	}
	size_t n = (size_t)c->cfg.n_envs * c->geoms.size();
	// Small batches (the reference's own case is ONE mjData): the whole call - one H2D copy, every kernel, the result
	// copies - is a CUDA graph captured on the second step after hcs_finalize and replayed from then on: one launch
	// instead of three copies and 4 - 30 kernel launches.  The inputs go through the context's pinned staging buffer.
	static const bool graphs_ok = getenv("HCS_NO_GRAPH") == nullptr;
	if (graphs_ok && c->h_in && !c->profiling) {
		const int key = with_sensors ? 1 : 0;
		memcpy(c->h_in, xpos, n * 3 * sizeof(double));
		memcpy(c->h_in + n * 3, xmat, n * 9 * sizeof(double));
		memcpy(c->h_in + n * 12, vel, n * 6 * sizeof(double));
		const bool direct = !with_sensors && c->dh_wrench && c->dh_flags;
		auto enqueue = [&]() {
			CK(cudaMemcpyAsync(c->d_xpos, c->h_in, n * 18 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
			step_device(c, c->d_xpos, c->d_xmat, c->d_vel, with_sensors, direct);
			if (!direct)
				fetch_enqueue(c, with_sensors, /*with_pairs=*/false);
		};
		const int set = next_counter_set(c); // the graph of this set: it bakes the counters' addresses in
		if (set >= 0 && c->step_graph[key][set]) {
			select_counter_set(c, set);
			c->set_clean[set] = false, c->set_clean[1 - set] = true;
			CK(cudaGraphLaunch(c->step_graph[key][set], c->stream));
			c->kernels_last_step = c->graph_kernels[key][set];
			c->step_counter++;
			c->last_with_sensors = with_sensors != 0;
		} else if (set >= 0 && c->steps_since_finalize >= 1) { // (the first step runs eagerly: one-time function attributes are set there)
			cudaGraph_t g = nullptr;
			CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
			try {
				enqueue();
			} catch (...) {
				cudaStreamEndCapture(c->stream, &g);
				if (g)
					cudaGraphDestroy(g);
				cudaGetLastError();
				throw;
			}
			CK(cudaStreamEndCapture(c->stream, &g));
			cudaError_t ie = cudaGraphInstantiate(&c->step_graph[key][set], g, 0);
			cudaGraphDestroy(g);
			if (ie != cudaSuccess) {
				cudaGetLastError();
				c->step_graph[key][set] = nullptr;
				c->set_clean[set]       = true; // (nothing ran: the capture only recorded the step)
				c->set_clean[1 - set]   = false;
				enqueue(); // no graph on this driver: plain launches
			} else {
				c->graph_kernels[key][set] = c->kernels_last_step;
				CK(cudaGraphLaunch(c->step_graph[key][set], c->stream));
			}
		} else {
			enqueue();
		}
		c->steps_since_finalize++;
		CK(cudaStreamSynchronize(c->stream));
		c->results_on_host = true;
		c->pairs_on_host   = false;
		c->sensors_on_host = with_sensors != 0;
		return check_flags(c);
	}
	// Inputs are staged with three DMA copies.  Reading the poses straight from pinned host memory inside the kernels
	// (zero-copy) was measured and dropped as the default: the reads are 8-byte uniform loads, which PCIe serves far
	// below its copy bandwidth (C1 +3 % end to end, C4 with five geoms -9 %); HCS_ZERO_COPY_IN=1 keeps the experiment.
	// Outputs: without sensors the finalize kernel is the last kernel of the step and writes the per-geom wrenches
	// and the flags into the context's mapped pinned mirrors itself (posted PCIe writes), so nothing is copied
	// after the kernels; with sensors the images are copied back behind the sensor kernels as before.
	const double *in[3] = { c->d_xpos, c->d_xmat, c->d_vel };
	static const bool zero_copy_in = getenv("HCS_ZERO_COPY_IN") != nullptr;
	bool staged = true;
	if (zero_copy_in && !with_sensors) {
		const double *host[3] = { xpos, xmat, vel }, *dev[3] = { nullptr, nullptr, nullptr };
		bool pinned = true;
		for (int k = 0; k < 3 && pinned; ++k) {
			cudaPointerAttributes attr{};
			if (cudaPointerGetAttributes(&attr, host[k]) != cudaSuccess) {
				cudaGetLastError();
				pinned = false;
			} else {
				pinned = attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
				dev[k] = (const double *)attr.devicePointer;
			}
		}
		if (pinned)
			in[0] = dev[0], in[1] = dev[1], in[2] = dev[2], staged = false;
	}
	if (staged) {
		CK(cudaMemcpyAsync(c->d_xpos, xpos, n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CK(cudaMemcpyAsync(c->d_xmat, xmat, n * 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CK(cudaMemcpyAsync(c->d_vel, vel, n * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	}
	static const bool direct_out_ok = getenv("HCS_NO_DIRECT_OUT") == nullptr;
	// (measured: 393 KB of wrenches on C1: +6 % end to end; 983 KB on C4: -3 %, the 8-byte posted writes lose to one
	// DMA burst once the output is large)
	const size_t wrench_bytes = n * 6 * sizeof(double);
	if (direct_out_ok && !with_sensors && c->dh_wrench && c->dh_flags && wrench_bytes <= (512u << 10)) {
		step_device(c, in[0], in[1], in[2], 0, /*direct_out=*/true);
		CK(cudaStreamSynchronize(c->stream));
		c->results_on_host = true; // wrenches and flags are in the pinned mirrors; pair results stay on the device
		return check_flags(c);
	}
	step_device(c, in[0], in[1], in[2], with_sensors);
	fetch(c, with_sensors, /*with_pairs=*/false);
	return check_flags(c);
	API_END(c)
}

// status of a finished pipelined step from its slot's flag mirror
static int slot_status(hcs_ctx *c, hcs_ctx::Slot &sl)
{
	int32_t keep[4];
	memcpy(keep, c->h_flags, sizeof keep);
	memcpy(c->h_flags, sl.h_flags, sizeof keep);
	int st = check_flags(c);
	memcpy(c->h_flags, keep, sizeof keep);
	return st;
}

int hcs_step_async(hcs_ctx *c, const double *xpos, const double *xmat, const double *vel, int with_sensors,
                   const hcs_outputs *out, int64_t *ticket)
{
	API_BEGIN(c)
	if (!c->finalized) {
		c->err = "hcs_step_async: call hcs_finalize first";
		return HCS_E_NOT_FINALIZED;
	}
	if (!xpos || !xmat || !vel || !ticket) {
		c->err = "hcs_step_async: null pose/velocity/ticket pointer";
		return HCS_E_INVALID;
	}
	const int64_t t     = c->next_ticket;
	hcs_ctx::Slot &sl   = c->slot[t % hcs_ctx::DEPTH];
	if (!sl.waited) { // the caller runs more than DEPTH steps ahead: finish the slot's previous step, keep its status
		CK(cudaEventSynchronize(sl.ev_done));
		sl.status = slot_status(c, sl);
		sl.waited = true;
	}
	const size_t n = (size_t)c->cfg.n_envs * c->geoms.size();
	const int n_env = c->cfg.n_envs;
	// copy-in stream: the slot's staging buffers are free once the kernels of the step that used them last are done
	if (sl.ticket >= 0)
		CK(cudaStreamWaitEvent(c->copy_in, sl.ev_free, 0));
	CK(cudaMemcpyAsync(sl.d_xpos, xpos, n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->copy_in));
	CK(cudaMemcpyAsync(sl.d_xmat, xmat, n * 9 * sizeof(double), cudaMemcpyHostToDevice, c->copy_in));
	CK(cudaMemcpyAsync(sl.d_vel, vel, n * 6 * sizeof(double), cudaMemcpyHostToDevice, c->copy_in));
	CK(cudaEventRecord(sl.ev_in, c->copy_in));
	// compute stream
	cudaStream_t s = c->stream;
	CK(cudaStreamWaitEvent(s, sl.ev_in, 0));
	if (sl.ticket >= 0) // the slot's wrench buffer must have left for the host before the finalize kernel rewrites it
		CK(cudaStreamWaitEvent(s, sl.ev_done, 0));
	memset(sl.h_flags, 0, 4 * sizeof(int32_t));
	const bool sensors = with_sensors != 0;
	step_device(c, sl.d_xpos, sl.d_xmat, sl.d_vel, with_sensors, false, sl.d_wrench, sensors ? nullptr : sl.dh_flags);
	CK(cudaEventRecord(sl.ev_free, s));
	if (sensors || !sl.dh_flags) {
		// sensor kernels follow the finalize kernel and raise flags of their own, and their device buffers exist once:
		// their results (and the flags) leave on the compute stream, in order
		CK(cudaMemcpyAsync(sl.h_flags, c->io.flags, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
		if (out && sensors) {
			for (size_t i = 0; i < c->sensors.size(); ++i)
				if (out->sensor_images && out->sensor_images[i])
					CK(cudaMemcpyAsync(out->sensor_images[i], c->sensors[i].dev.image,
					                   (size_t)n_env * c->sensors[i].cx * c->sensors[i].cy * sizeof(float), cudaMemcpyDeviceToHost, s));
			for (size_t i = 0; i < c->curved.size(); ++i)
				if (out->curved_values && out->curved_values[i])
					CK(cudaMemcpyAsync(out->curved_values[i], c->curved[i].dev.values,
					                   (size_t)n_env * c->curved[i].n_taxels() * sizeof(float), cudaMemcpyDeviceToHost, s));
			for (size_t i = 0; i < c->taxel.size(); ++i)
				if (out->taxel_values && out->taxel_values[i])
					CK(cudaMemcpyAsync(out->taxel_values[i], c->taxel[i].dev.values,
					                   (size_t)n_env * c->taxel[i].n_taxels() * sizeof(float), cudaMemcpyDeviceToHost, s));
		}
	}
	if (out && out->pair_results) // diagnostics: one device buffer, so in order on the compute stream
		CK(cudaMemcpyAsync(out->pair_results, c->io.pair_out, (size_t)n_env * c->pairs.size() * sizeof(hcs_pair_result),
		                   cudaMemcpyDeviceToHost, s));
	CK(cudaEventRecord(sl.ev_kdone, s));
	// copy-out stream: the wrenches of this step go to the caller while the next step's kernels run
	CK(cudaStreamWaitEvent(c->copy_out, sl.ev_kdone, 0));
	if (out && out->geom_wrench)
		CK(cudaMemcpyAsync(out->geom_wrench, sl.d_wrench, n * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->copy_out));
	CK(cudaEventRecord(sl.ev_done, c->copy_out));
	sl.ticket = t, sl.waited = false, sl.status = HCS_OK;
	c->next_ticket = t + 1;
	*ticket        = t;
	// the synchronous getters read the context's own buffers: the pipelined step left its wrenches in the slot
	c->results_on_host = c->pairs_on_host = c->sensors_on_host = false;
	return HCS_OK;
	API_END(c)
}

int hcs_wait(hcs_ctx *c, int64_t ticket)
{
	API_BEGIN(c)
	if (ticket < 0 || ticket >= c->next_ticket) {
		c->err = "hcs_wait: unknown ticket";
		return HCS_E_INVALID;
	}
	hcs_ctx::Slot &sl = c->slot[ticket % hcs_ctx::DEPTH];
	if (sl.ticket != ticket) {
		c->err = "hcs_wait: the ticket's results have been overwritten (more than 2 steps were in flight)";
		return HCS_E_INVALID;
	}
	if (!sl.waited) {
		CK(cudaEventSynchronize(sl.ev_done));
		sl.status = slot_status(c, sl);
		sl.waited = true;
	}
	if (sl.status != HCS_OK)
		(void)slot_status(c, sl); // sets the error text again
	return sl.status;
	API_END(c)
}

int hcs_sync(hcs_ctx *c)
{
	API_BEGIN(c)
	CK(cudaMemcpyAsync(c->h_flags, c->io.flags, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (c->profiling) {
		for (int i = 0; i < 5; ++i)
			cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]);
		c->stage_ms[5] = 0;
		cudaEventElapsedTime(&c->stage_ms[6], c->ev[0], c->ev[5]);
	}
	return check_flags(c);
	API_END(c)
}

int hcs_fetch_results(hcs_ctx *c, int with_sensors)
{
	API_BEGIN(c)
	if (!c->finalized)
		return HCS_E_NOT_FINALIZED;
	fetch(c, with_sensors && c->last_with_sensors);
	return check_flags(c);
	API_END(c)
}

int hcs_n_geoms(const hcs_ctx *c) { return c ? (int)c->geoms.size() : HCS_E_INVALID; }
int hcs_n_pairs(const hcs_ctx *c) { return c ? (int)c->pairs.size() : HCS_E_INVALID; }

int hcs_get_pair_results(hcs_ctx *c, hcs_pair_result *out)
{
	API_BEGIN(c)
	if (!c->finalized || !out)
		return HCS_E_INVALID;
	if (!c->pairs_on_host)
		fetch(c, c->sensors_on_host ? 1 : 0);
	memcpy(out, c->h_pair, (size_t)c->cfg.n_envs * c->pairs.size() * sizeof(hcs_pair_result));
	return HCS_OK;
	API_END(c)
}

int hcs_get_geom_wrenches(hcs_ctx *c, double *out)
{
	API_BEGIN(c)
	if (!c->finalized || !out)
		return HCS_E_INVALID;
	if (!c->results_on_host)
		fetch(c, 0);
	memcpy(out, c->h_wrench, (size_t)c->cfg.n_envs * c->geoms.size() * 6 * sizeof(double));
	return HCS_OK;
	API_END(c)
}

int hcs_get_sensor_image(hcs_ctx *c, int sensor, float *out)
{
	API_BEGIN(c)
	if (!c->finalized || !out || sensor < 0 || sensor >= (int)c->sensors.size())
		return HCS_E_INVALID;
	if (!c->last_with_sensors) {
		c->err = "hcs_get_sensor_image: the last step ran with with_sensors == 0";
		return HCS_E_INVALID;
	}
	if (!c->sensors_on_host)
		fetch(c, 1);
	SensorHost &s = c->sensors[sensor];
	memcpy(out, s.h_image, (size_t)c->cfg.n_envs * s.cx * s.cy * sizeof(float));
	return HCS_OK;
	API_END(c)
}

const hcs_pair_result *hcs_device_pair_results(hcs_ctx *c) { return c && c->finalized ? c->io.pair_out : nullptr; }
const double *hcs_device_geom_wrenches(hcs_ctx *c) { return c && c->finalized ? c->io.geom_wrench : nullptr; }
const float *hcs_device_sensor_image(hcs_ctx *c, int sensor)
{
	if (!c || !c->finalized || sensor < 0 || sensor >= (int)c->sensors.size())
		return nullptr;
	return c->sensors[sensor].dev.image;
}

int hcs_get_faces(hcs_ctx *c, hcs_face *out, int cap)
{
	API_BEGIN(c)
	if (!c->finalized || c->io.max_faces <= 0) {
		c->err = "hcs_get_faces: needs hcs_config.max_faces > 0";
		return HCS_E_INVALID;
	}
	int32_t n = 0;
	CK(cudaMemcpyAsync(&n, c->io.face_count, sizeof n, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (n > c->io.max_faces) {
		c->err = "per-face dump overflow: raise hcs_config.max_faces";
		return HCS_E_CAPACITY;
	}
	int m = std::min<int>(n, cap);
	if (m > 0 && out) {
		CK(cudaMemcpyAsync(out, c->io.faces, (size_t)m * sizeof(hcs_face), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return n;
	API_END(c)
}

int hcs_get_face_vertices(hcs_ctx *c, double *out, int cap)
{
	API_BEGIN(c)
	if (!c->finalized || !c->io.face_verts) {
		c->err = "hcs_get_face_vertices: needs hcs_config.max_faces > 0 and hcs_config.face_vertices";
		return HCS_E_INVALID;
	}
	int32_t n = 0;
	CK(cudaMemcpyAsync(&n, c->io.face_count, sizeof n, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (n > c->io.max_faces) {
		c->err = "per-face dump overflow: raise hcs_config.max_faces";
		return HCS_E_CAPACITY;
	}
	int m = std::min<int>(n, cap);
	if (m > 0 && out) {
		CK(cudaMemcpyAsync(out, c->io.face_verts, (size_t)m * HCS_FACE_VERTEX_STRIDE * sizeof(double),
		                   cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return n;
	API_END(c)
}

int hcs_get_emitted(hcs_ctx *c, int env, int pair, int32_t *out, int cap)
{
	API_BEGIN(c)
	if (!c->finalized || env < 0 || env >= c->cfg.n_envs || pair < 0 || pair >= (int)c->pairs.size())
		return HCS_E_INVALID;
	const PairDesc &P = c->pair_desc[pair];
	if (P.kind == PAIR_NONE)
		return 0;
	// The pair's flat candidate list of the last step holds the whole batch in no particular order: it is copied once
	// per step and indexed by environment, so that asking for every environment in turn costs one copy.
	if (c->emitted_cache.size() != c->pairs.size())
		c->emitted_cache.assign(c->pairs.size(), EmittedCache());
	EmittedCache &E = c->emitted_cache[pair];
	if (E.step != c->step_counter) {
		int32_t used[PAIR_COUNTERS] = { 0, 0, 0, 0 };
		CK(cudaMemcpyAsync(used, P.counters, sizeof used, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		const int total = std::min(std::max(used[0], 0), P.contrib_cap);
		std::vector<uint4> flat(total);
		std::vector<uint8_t> nv(total);
		if (total > 0) {
			CK(cudaMemcpyAsync(flat.data(), P.flat, (size_t)total * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
			CK(cudaMemcpyAsync(nv.data(), P.nverts, (size_t)total, cudaMemcpyDeviceToHost, c->stream));
			CK(cudaStreamSynchronize(c->stream));
		}
		E.offset.assign((size_t)c->cfg.n_envs + 1, 0);
		for (int i = 0; i < total; ++i)
			if ((nv[i] & 15) >= 3 && (int)flat[i].z < c->cfg.n_envs)
				E.offset[flat[i].z + 1]++;
		for (int e = 0; e < c->cfg.n_envs; ++e)
			E.offset[e + 1] += E.offset[e];
		E.triples.assign(3 * (size_t)E.offset[c->cfg.n_envs], 0);
		std::vector<int> cur(E.offset.begin(), E.offset.end() - 1);
		for (int i = 0; i < total; ++i)
			if ((nv[i] & 15) >= 3 && (int)flat[i].z < c->cfg.n_envs) {
				const int eA = (int)(flat[i].y & CAND_ELEM_MASK), eB = (int)flat[i].x; // (tree element of A, query element of B)
				int32_t *o = &E.triples[3 * (size_t)cur[flat[i].z]++];
				o[0] = P.sign > 0 ? eA : eB, o[1] = P.sign > 0 ? eB : eA, o[2] = nv[i] & 15;
			}
		E.step = c->step_counter;
	}
	const int n = E.offset[env + 1] - E.offset[env];
	if (out && cap > 0 && n > 0)
		memcpy(out, &E.triples[3 * (size_t)E.offset[env]], (size_t)std::min(n, cap) * 3 * sizeof(int32_t));
	return n;
	API_END(c)
}

// the environment's tactile triangles in canonical order (pair, query element, tree element, fan triangle)
static std::vector<TactileTri> env_triangles(hcs_ctx *c, int env)
{
	std::vector<TactileTri> mine;
	if (c->io.max_tris <= 0)
		return mine;
	int32_t n = 0;
	CK(cudaMemcpyAsync(&n, c->io.tri_count, sizeof n, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	n = std::min(n, c->io.max_tris);
	std::vector<TactileTri> pool(n);
	if (n > 0) {
		CK(cudaMemcpyAsync(pool.data(), c->io.tri_pool, (size_t)n * sizeof(TactileTri), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	for (const TactileTri &t : pool)
		if (t.env == env)
			mine.push_back(t);
	std::sort(mine.begin(), mine.end(), [](const TactileTri &a, const TactileTri &b) {
		return a.key_hi != b.key_hi ? a.key_hi < b.key_hi : a.key_lo < b.key_lo;
	});
	return mine;
}

int hcs_get_tactile_triangles(hcs_ctx *c, int env, double *out, int cap)
{
	API_BEGIN(c)
	if (!c->finalized || env < 0 || env >= c->cfg.n_envs)
		return HCS_E_INVALID;
	const std::vector<TactileTri> mine = env_triangles(c, env);
	int m = 0;
	for (const TactileTri &t : mine) {
		if (m < cap && out) {
			for (int k = 0; k < 9; ++k)
				out[12 * m + k] = t.v[k];
			for (int k = 0; k < 3; ++k)
				out[12 * m + 9 + k] = t.e[k];
		}
		++m;
	}
	return m;
	API_END(c)
}

int hcs_get_tactile_triangle_pairs(hcs_ctx *c, int env, int32_t *pair_out, int cap)
{
	API_BEGIN(c)
	if (!c->finalized || env < 0 || env >= c->cfg.n_envs)
		return HCS_E_INVALID;
	const std::vector<TactileTri> mine = env_triangles(c, env);
	int m = 0;
	for (const TactileTri &t : mine) {
		if (m < cap && pair_out)
			pair_out[m] = (int32_t)(t.key_hi >> TRI_PAIR_SHIFT);
		++m;
	}
	return m;
	API_END(c)
}

int hcs_geom_info(const hcs_ctx *c, int geom, int info[3])
{
	if (!c || geom < 0 || geom >= (int)c->geoms.size() || !info)
		return HCS_E_INVALID;
	const HostMesh &m = c->geoms[geom].mesh;
	info[0]           = m.plane ? 2 : (m.soft ? 1 : 0);
	info[1]           = m.n_verts();
	info[2]           = m.n_elems();
	return HCS_OK;
}

int hcs_get_mesh(hcs_ctx *c, int geom, double *verts, int32_t *elems, double *pressure, double *grad, double *e0)
{
	API_BEGIN(c)
	if (!c->finalized || geom < 0 || geom >= (int)c->geoms.size()) {
		c->err = "hcs_get_mesh: finalize first";
		return HCS_E_INVALID;
	}
	const GeomDev &d = c->geoms[geom].dev;
	if (d.kind == 2)
		return HCS_OK;
	cudaStream_t s = c->stream;
	if (verts)
		CK(cudaMemcpyAsync(verts, d.verts, (size_t)d.n_verts * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
	if (elems)
		CK(cudaMemcpyAsync(elems, d.elems, (size_t)d.n_elems * (d.kind == 1 ? 4 : 3) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
	if (d.kind == 1) {
		if (pressure)
			CK(cudaMemcpyAsync(pressure, d.pressure, (size_t)d.n_verts * sizeof(double), cudaMemcpyDeviceToHost, s));
		std::vector<TetField> tf(d.n_elems);
		CK(cudaMemcpyAsync(tf.data(), d.tet_field, (size_t)d.n_elems * sizeof(TetField), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		for (int t = 0; t < d.n_elems; ++t) {
			if (grad)
				for (int k = 0; k < 3; ++k)
					grad[3 * (size_t)t + k] = tf[t].grad[k];
			if (e0)
				e0[t] = tf[t].e0;
		}
	} else {
		std::vector<TriRec> tr(d.n_elems);
		CK(cudaMemcpyAsync(tr.data(), d.tris, (size_t)d.n_elems * sizeof(TriRec), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		if (grad)
			for (int t = 0; t < d.n_elems; ++t)
				for (int k = 0; k < 3; ++k)
					grad[3 * (size_t)t + k] = tr[t].n[k];
	}
	CK(cudaStreamSynchronize(s));
	return HCS_OK;
	API_END(c)
}

int hcs_get_lbvh(hcs_ctx *c, int geom, void *out_nodes, int max_nodes)
{
	API_BEGIN(c)
	if (!c->finalized || geom < 0 || geom >= (int)c->geoms.size() || c->geoms[geom].dev.kind != 1) {
		c->err = "hcs_get_lbvh: needs a finalized context and a soft geom";
		return HCS_E_INVALID;
	}
	const GeomDev &d = c->geoms[geom].dev;
	int n            = std::max(1, d.n_elems - 1);
	if (out_nodes && max_nodes > 0) {
		CK(cudaMemcpyAsync(out_nodes, d.nodes, (size_t)std::min(n, max_nodes) * sizeof(BvhNode), cudaMemcpyDeviceToHost,
		                   c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return n;
	API_END(c)
}

int hcs_get_counters(hcs_ctx *c, int64_t out[5])
{
	API_BEGIN(c)
	if (!c->finalized || !out)
		return HCS_E_INVALID;
	if (!c->pairs_on_host)
		fetch(c, c->sensors_on_host ? 1 : 0);
	int64_t cand = 0, poly = 0, faces = 0;
	size_t n = (size_t)c->cfg.n_envs * c->pairs.size();
	for (size_t i = 0; i < n; ++i) {
		cand += c->h_pair[i].n_candidates;
		poly += c->h_pair[i].n_polygons;
		faces += c->h_pair[i].n_faces;
	}
	int32_t ntri = 0;
	if (c->io.max_tris > 0) {
		CK(cudaMemcpyAsync(&ntri, c->io.tri_count, sizeof ntri, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	out[0] = cand, out[1] = poly, out[2] = faces, out[3] = ntri, out[4] = c->kernels_last_step;
	return HCS_OK;
	API_END(c)
}

int hcs_set_profiling(hcs_ctx *c, int enable)
{
	if (!c)
		return HCS_E_INVALID;
	c->profiling = enable != 0;
	return HCS_OK;
}

int hcs_get_stage_ms(hcs_ctx *c, float out[7])
{
	if (!c || !out)
		return HCS_E_INVALID;
	for (int i = 0; i < 7; ++i)
		out[i] = c->stage_ms[i];
	return HCS_OK;
}

} // extern "C"
