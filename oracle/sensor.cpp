// CPU ORACLE (test infrastructure; ray caster pinned to oracle/_ref, the rest unpinned: oracle.hpp) — Myrmex flat tactile sensor.
//
// Restates mujoco_contact_surface_sensors/src/flat_tactile_sensor.cpp:127-214 (load constants),
// :262-402 (bvh_update) and the float32 ray caster of src/bvh.cpp:49-476 +
// include/mujoco_contact_surface_sensors/bvh.h:157-176 (scalar, non-SSE code path).
// The mixed float/double arithmetic of the reference is reproduced operation by operation
// (SURVEY.md App. B.3/B.4).  Deviation noted in DESIGN.md (Q16): mju_rotVecMat(pos,pos,rot) is
// evaluated alias-safe (the mathematically intended rotation).
#include "oracle.hpp"

#include <array>
#include <random>

#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace orc {

namespace {

struct F3 {
	float x, y, z;
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline F3 operator-(F3 a, F3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline F3 operator+(F3 a, F3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline F3 operator*(F3 a, float b) { return { a.x * b, a.y * b, a.z * b }; }
inline F3 operator-(F3 a) { return { -a.x, -a.y, -a.z }; }
inline F3 crossf(F3 a, F3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float dotf(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // float3.h: left to right
inline float fmin_(float a, float b) { return a < b ? a : b; }
inline float fmax_(float a, float b) { return a > b ? a : b; }
inline F3 fmin3(F3 a, F3 b) { return { fmin_(a.x, b.x), fmin_(a.y, b.y), std::fmin(a.z, b.z) }; }
inline F3 fmax3(F3 a, F3 b) { return { fmax_(a.x, b.x), fmax_(a.y, b.y), std::fmax(a.z, b.z) }; }

struct Tri {
	F3 v0, v1, v2, centroid;
};
struct Ray {
	F3 O, D, rD;
	float t, u, v;
	unsigned hit; // (blas << 20) + triangle
};

// bvh.cpp:49-74
inline void intersect_triangle(Ray &ray, const Tri &tri, unsigned id)
{
	F3 edge1 = tri.v1 - tri.v0, edge2 = tri.v2 - tri.v0;
	F3 h    = crossf(ray.D, edge2);
	float a = dotf(edge1, h);
	if (std::fabs(a) < 1e-10f)
		return;
	float f = 1.0f / a;
	F3 s    = ray.O - tri.v0;
	float u = f * dotf(s, h);
	if (u < 0.0f || u > 1.0f)
		return;
	F3 q    = crossf(s, edge1);
	float v = f * dotf(ray.D, q);
	if (v < 0.0f || u + v > 1.0f)
		return;
	float t = f * dotf(edge2, q);
	if (t < ray.t && t > 0.0f) {
		ray.t   = t;
		ray.u   = u;
		ray.v   = v;
		ray.hit = id;
	}
}

// bvh.h:157-176
inline float intersect_aabb(const Ray &ray, F3 bmin, F3 bmax)
{
	float tx1 = (bmin.x - ray.O.x) * ray.rD.x, tx2 = (bmax.x - ray.O.x) * ray.rD.x;
	float tmin = std::min(tx1, tx2), tmax = std::max(tx1, tx2);
	float ty1 = (bmin.y - ray.O.y) * ray.rD.y, ty2 = (bmax.y - ray.O.y) * ray.rD.y;
	tmin      = std::max(tmin, std::min(ty1, ty2));
	tmax      = std::min(tmax, std::max(ty1, ty2));
	float tz1 = (bmin.z - ray.O.z) * ray.rD.z, tz2 = (bmax.z - ray.O.z) * ray.rD.z;
	tmin      = std::max(tmin, std::min(tz1, tz2));
	tmax      = std::min(tmax, std::max(tz1, tz2));
	if (tmax >= tmin && tmin < ray.t && tmax > 0)
		return tmin;
	return 1e30f;
}

struct Node {
	F3 mn;
	unsigned left_first;
	F3 mx;
	unsigned count;
};
inline float area_of(F3 mn, F3 mx)
{
	F3 s = mx - mn;
	return s.x * s.y + s.x * s.z + s.y * s.z;
}

const int BINS = 8;

// bvh.cpp:90-357: binned-SAH BLAS over the float32 copy of one contact surface
struct Blas {
	std::vector<Tri> tri;
	std::vector<unsigned> idx;
	std::vector<Node> node;
	unsigned used = 2;
	F3 bmin, bmax;
	const Surface *surface = nullptr;

	void update_bounds(unsigned ni, F3 &cmin, F3 &cmax)
	{
		Node &n = node[ni];
		n.mn    = { 1e30f, 1e30f, 1e30f };
		n.mx    = { -1e30f, -1e30f, -1e30f };
		cmin    = n.mn;
		cmax    = n.mx;
		for (unsigned i = 0; i < n.count; ++i) {
			const Tri &t = tri[idx[n.left_first + i]];
			n.mn         = fmin3(fmin3(fmin3(n.mn, t.v0), t.v1), t.v2);
			n.mx         = fmax3(fmax3(fmax3(n.mx, t.v0), t.v1), t.v2);
			cmin         = fmin3(cmin, t.centroid);
			cmax         = fmax3(cmax, t.centroid);
		}
	}
	float best_split(const Node &n, int &axis, int &split, F3 cmin, F3 cmax)
	{
		float best = 1e30f;
		for (int a = 0; a < 3; ++a) {
			float bmn = cmin[a], bmx = cmax[a];
			float scale = BINS / (bmx - bmn);
			struct Bin {
				F3 mn{ 1e30f, 1e30f, 1e30f }, mx{ -1e30f, -1e30f, -1e30f };
				int count = 0;
			} bin[BINS];
			for (unsigned i = 0; i < n.count; ++i) {
				const Tri &t = tri[idx[n.left_first + i]];
				int b        = std::max(std::min(BINS - 1, (int)((t.centroid[a] - bmn) * scale)), 0);
				bin[b].count++;
				bin[b].mn = fmin3(fmin3(fmin3(bin[b].mn, t.v0), t.v1), t.v2);
				bin[b].mx = fmax3(fmax3(fmax3(bin[b].mx, t.v0), t.v1), t.v2);
			}
			float la[BINS - 1], ra[BINS - 1];
			F3 lmn{ 1e30f, 1e30f, 1e30f }, lmx{ -1e30f, -1e30f, -1e30f }, rmn = lmn, rmx = lmx;
			int ls = 0, rs = 0;
			for (int i = 0; i < BINS - 1; ++i) {
				ls += bin[i].count;
				if (bin[i].mn.x != 1e30f) {
					lmn = fmin3(fmin3(lmn, bin[i].mn), bin[i].mx);
					lmx = fmax3(fmax3(lmx, bin[i].mn), bin[i].mx);
				}
				la[i] = area_of(lmn, lmx) * ls;
				rs += bin[BINS - 1 - i].count;
				if (bin[BINS - 1 - i].mn.x != 1e30f) {
					rmn = fmin3(fmin3(rmn, bin[BINS - 1 - i].mn), bin[BINS - 1 - i].mx);
					rmx = fmax3(fmax3(rmx, bin[BINS - 1 - i].mn), bin[BINS - 1 - i].mx);
				}
				ra[BINS - 2 - i] = area_of(rmn, rmx) * rs;
			}
			for (int i = 0; i < BINS - 1; ++i) {
				float cost = la[i] + ra[i];
				if (cost < best) {
					best  = cost;
					axis  = a;
					split = i + 1;
				}
			}
		}
		return best;
	}
	void subdivide(unsigned ni, F3 cmin, F3 cmax)
	{
		int axis = 0, split = 0;
		float cost   = best_split(node[ni], axis, split, cmin, cmax);
		float nosplit = area_of(node[ni].mn, node[ni].mx) * node[ni].count;
		if (cost >= nosplit)
			return;
		int i = node[ni].left_first, j = i + node[ni].count - 1;
		float scale = BINS / (cmax[axis] - cmin[axis]);
		while (i <= j) {
			int b = std::min(BINS - 1, (int)((tri[idx[i]].centroid[axis] - cmin[axis]) * scale));
			if (b < split)
				i++;
			else
				std::swap(idx[i], idx[j--]);
		}
		unsigned left_count = i - node[ni].left_first;
		if (left_count == 0 || left_count == node[ni].count)
			return;
		unsigned l = used++, r = used++;
		node[l].left_first  = node[ni].left_first;
		node[l].count       = left_count;
		node[r].left_first  = i;
		node[r].count       = node[ni].count - left_count;
		node[ni].left_first = l;
		node[ni].count      = 0;
		F3 cmn, cmx;
		update_bounds(l, cmn, cmx);
		subdivide(l, cmn, cmx);
		update_bounds(r, cmn, cmx);
		subdivide(r, cmn, cmx);
	}
	void build(const Surface &s)
	{
		surface = &s;
		int n   = s.num_faces();
		tri.resize(n);
		idx.resize(n);
		for (int i = 0; i < n; ++i) {
			const int *f = &s.face_idx[s.face_first[i]];
			auto cvt     = [&](int v) { return F3{ (float)s.v[v].x, (float)s.v[v].y, (float)s.v[v].z }; };
			tri[i].v0 = cvt(f[0]);
			tri[i].v1 = cvt(f[1]);
			tri[i].v2 = cvt(f[2]);
			tri[i].centroid = ((tri[i].v0 + tri[i].v1) + tri[i].v2) * 0.3333f;
			idx[i]          = i;
		}
		node.assign(2 * std::max(n, 1), Node{ { 0, 0, 0 }, 0, { 0, 0, 0 }, 0 });
		used               = 2;
		node[0].left_first = 0;
		node[0].count      = n;
		F3 cmn, cmx;
		update_bounds(0, cmn, cmx);
		subdivide(0, cmn, cmx);
		bmin = node[0].mn;
		bmax = node[0].mx;
	}
	void intersect(Ray &ray, unsigned blas_idx) const
	{
		const Node *n = &node[0], *stack[64];
		unsigned sp   = 0;
		while (true) {
			if (n->count > 0) {
				for (unsigned i = 0; i < n->count; ++i)
					intersect_triangle(ray, tri[idx[n->left_first + i]], (blas_idx << 20) + idx[n->left_first + i]);
				if (sp == 0)
					break;
				n = stack[--sp];
				continue;
			}
			const Node *c1 = &node[n->left_first], *c2 = &node[n->left_first + 1];
			float d1 = intersect_aabb(ray, c1->mn, c1->mx), d2 = intersect_aabb(ray, c2->mn, c2->mx);
			if (d1 > d2) {
				std::swap(c1, c2);
				std::swap(d1, d2);
			}
			if (d1 == 1e30f) {
				if (sp == 0)
					break;
				n = stack[--sp];
			} else {
				n = c1;
				if (d2 != 1e30f)
					stack[sp++] = c2;
			}
		}
	}
};

// bvh.cpp:359-476: agglomerative TLAS over the BLAS bounds
struct Tlas {
	struct TNode {
		F3 mn;
		unsigned left_right;
		F3 mx;
		unsigned blas;
	};
	std::vector<TNode> node;
	const std::vector<Blas> *blas = nullptr;

	void build(const std::vector<Blas> &b)
	{
		blas  = &b;
		int n = (int)b.size();
		node.assign(2 * n + 2, TNode{ { 0, 0, 0 }, 0, { 0, 0, 0 }, 0 });
		std::vector<unsigned> ni(n);
		unsigned used = 1;
		for (int i = 0; i < n; ++i) {
			ni[i]                 = used;
			node[used].mn         = b[i].bmin;
			node[used].mx         = b[i].bmax;
			node[used].blas       = i;
			node[used].left_right = 0;
			used++;
		}
		auto best_match = [&](int N, int A) {
			float smallest = 1e30f;
			int best       = -1;
			for (int B = 0; B < N; ++B)
				if (B != A) {
					F3 mx = fmax3(node[ni[A]].mx, node[ni[B]].mx), mn = fmin3(node[ni[A]].mn, node[ni[B]].mn);
					float sa = area_of(mn, mx);
					if (sa < smallest)
						smallest = sa, best = B;
				}
			return best;
		};
		int cnt = n, A = 0, B = n > 1 ? best_match(cnt, A) : -1;
		while (cnt > 1) {
			int C = best_match(cnt, B);
			if (A == C) {
				unsigned ia = ni[A], ib = ni[B];
				node[used].left_right = ia + (ib << 16);
				node[used].mn         = fmin3(node[ia].mn, node[ib].mn);
				node[used].mx         = fmax3(node[ia].mx, node[ib].mx);
				ni[A]                 = used++;
				ni[B]                 = ni[cnt - 1];
				B                     = best_match(--cnt, A);
			} else {
				A = B;
				B = C;
			}
		}
		node[0] = node[ni[A]];
	}
	void intersect(Ray &ray) const
	{
		ray.rD = { 1.0f / ray.D.x, 1.0f / ray.D.y, 1.0f / ray.D.z };
		const TNode *n = &node[0], *stack[64];
		unsigned sp    = 0;
		while (true) {
			if (n->left_right == 0) {
				(*blas)[n->blas].intersect(ray, n->blas);
				if (sp == 0)
					break;
				n = stack[--sp];
				continue;
			}
			const TNode *c1 = &node[n->left_right & 0xffff], *c2 = &node[n->left_right >> 16];
			float d1 = intersect_aabb(ray, c1->mn, c1->mx), d2 = intersect_aabb(ray, c2->mn, c2->mx);
			if (d1 > d2) {
				std::swap(c1, c2);
				std::swap(d1, d2);
			}
			if (d1 == 1e30f) {
				if (sp == 0)
					break;
				n = stack[--sp];
			} else {
				n = c1;
				if (d2 != 1e30f)
					stack[sp++] = c2;
			}
		}
	}
};

const float SQRT_2f = 1.41421356237; // flat_tactile_sensor.h:47 (float)

ExternalCaster g_ext{ nullptr, nullptr, nullptr };

} // namespace

// caster == 2: the same surfaces handed to the REFERENCE's own BVH/TLAS (oracle/_ref, compiled from
// mujoco_contact_surface_sensors/src/bvh.cpp) as triangle soups in double, the way bvh.cpp:92-105 reads them
static void *make_external_tlas(const std::vector<Blas> &blas)
{
	if (!g_ext.create)
		throw std::runtime_error("no external ray caster installed (oracle/_ref not loaded)");
	std::vector<int> n_tri;
	std::vector<double> soup, press;
	for (const Blas &b : blas) {
		const Surface &s = *b.surface;
		n_tri.push_back(s.num_faces());
		for (int f = 0; f < s.num_faces(); ++f)
			for (int k = 0; k < 3; ++k) {
				int v = s.face_idx[s.face_first[f] + k];
				soup.push_back(s.v[v].x), soup.push_back(s.v[v].y), soup.push_back(s.v[v].z);
				press.push_back(s.e[v]);
			}
	}
	return g_ext.create((int)blas.size(), n_tri.data(), soup.data(), press.data());
}
static void external_cast(void *ext, Ray &ray)
{
	float O[3] = { ray.O.x, ray.O.y, ray.O.z }, D[3] = { ray.D.x, ray.D.y, ray.D.z }, tuv[3];
	uint32_t hid;
	g_ext.cast(ext, 1, O, D, tuv, &hid);
	ray.t = tuv[0], ray.u = tuv[1], ray.v = tuv[2], ray.hit = hid;
}

// flat_tactile_sensor.cpp:262-402 bvh_update.  use_bvh=false replaces the TLAS/BLAS traversal by a
// linear scan over all triangles in surface order (same Möller–Trumbore arithmetic, same strict
// `t < hit.t` rule) and is used to bound tie-breaking effects.
void flat_sensor_image(const Scene &sc, const StepState &st, int sensor, float *out, int caster, bool parallel,
                       FlatTrace *trace)
{
	const bool use_bvh = caster == 1;
	const bool use_ext = caster == 2;

	const FlatSensor &fs = sc.sensors[sensor];
	int id = fs.geom, cx = fs.cx, cy = fs.cy, S = fs.S;
	float xs = (float)fs.size[0], ys = (float)fs.size[1], zs = (float)fs.size[2];
	double resolution = fs.resolution;
	// load(): :179-185
	float di_factor     = resolution / S;
	float sub_halfwidth = di_factor / 2.0f - resolution / 2.0f;
	float rmean         = 1. / (S * S);
	float rS            = resolution / S;
	float max_dist      = SQRT_2f * resolution / 2.0f;
	float sigma         = fs.sigma;
	bool use_gaussian = fs.window == 1, use_tukey = fs.window == 2, use_square = fs.window == 3;

	double rot[9];
	std::memcpy(rot, &st.xmat[9 * id], sizeof(rot));
	F3 sensor_normal{ (float)rot[2], (float)rot[5], (float)rot[8] };
	double sensor_xpos[3]    = { st.xpos[3 * id], st.xpos[3 * id + 1], st.xpos[3 * id + 2] };
	double sensor_topleft[3] = { -xs, -ys, zs };

	std::fill(out, out + cx * cy, 0.0f);
	std::vector<Blas> blas;
	for (const PairOut &po : st.out)
		if (po.has_surface && (po.gM == id || po.gN == id) && po.s->tri)
			blas.emplace_back();
	{
		int b = 0;
		for (const PairOut &po : st.out)
			if (po.has_surface && (po.gM == id || po.gN == id) && po.s->tri)
				blas[b++].build(*po.s);
	}
	if (blas.empty())
		return;
	Tlas tlas;
	tlas.build(blas);
	void *ext = use_ext ? make_external_tlas(blas) : nullptr;
	if (trace) {
		trace->rays.assign((size_t)cx * cy * S * S * 6, 0.0f);
		trace->tuv.assign((size_t)cx * cy * S * S * 3, 0.0f);
		trace->id.assign((size_t)cx * cy * S * S, 0u);
	}

	float rsigma_squared = 0, rtukey = 0;
	if (use_gaussian)
		rsigma_squared = 0.5f / (sigma * sigma);
	else if (use_tukey)
		rtukey = 1.f / sigma * S * S;

#pragma omp parallel for collapse(2) schedule(dynamic) if (parallel)
	for (int x = 0; x < cx; x++) {
		for (int y = 0; y < cy; y++) {
			float avg_pressure = 0;
			for (int i = 0; i < S; i++) {
				for (int j = 0; j < S; j++) {
					double pos[3] = { sensor_topleft[0] + x * resolution + i * rS + 0.5 * rS,
						              sensor_topleft[1] + y * resolution + j * rS + 0.5 * rS, 1.5 * zs };
					double w[3]   = { rot[0] * pos[0] + rot[1] * pos[1] + rot[2] * pos[2],
						              rot[3] * pos[0] + rot[4] * pos[1] + rot[5] * pos[2],
						              rot[6] * pos[0] + rot[7] * pos[1] + rot[8] * pos[2] };
					w[0] += sensor_xpos[0];
					w[1] += sensor_xpos[1];
					w[2] += sensor_xpos[2];
					F3 sensor_point{ (float)w[0], (float)w[1], (float)w[2] };
					Ray ray;
					ray.O   = sensor_point + sensor_normal * (float)1e-8;
					ray.D   = -sensor_normal;
					ray.t   = 1e30f;
					ray.u = ray.v = 0;
					ray.hit = 0;
					if (use_ext) {
						external_cast(ext, ray);
					} else if (use_bvh) {
						tlas.intersect(ray);
					} else {
						ray.rD = { 1.0f / ray.D.x, 1.0f / ray.D.y, 1.0f / ray.D.z };
						for (unsigned b = 0; b < blas.size(); ++b)
							for (unsigned t = 0; t < blas[b].tri.size(); ++t)
								intersect_triangle(ray, blas[b].tri[t], (b << 20) + t);
					}
					if (trace) {
						size_t r = (((size_t)x * cy + y) * S + i) * S + j;
						float *q = &trace->rays[6 * r];
						q[0] = ray.O.x, q[1] = ray.O.y, q[2] = ray.O.z, q[3] = ray.D.x, q[4] = ray.D.y, q[5] = ray.D.z;
						trace->tuv[3 * r] = ray.t, trace->tuv[3 * r + 1] = ray.u, trace->tuv[3 * r + 2] = ray.v;
						trace->id[r] = ray.hit;
					}
					if (ray.t < 1.5 * zs && ray.t > 0.0f) {
						double bary[3]   = { (double)(1 - ray.u - ray.v), (double)ray.u, (double)ray.v };
						unsigned tri_idx = ray.hit & 0xFFFFF, blas_idx = ray.hit >> 20;
						const Surface &s = *blas[blas_idx].surface;
						const int *f     = &s.face_idx[s.face_first[tri_idx]];
						double ev        = bary[0] * s.e[f[0]];
						ev += bary[1] * s.e[f[1]];
						ev += bary[2] * s.e[f[2]];
						float raw    = ev * rmean;
						float weight = 1.0f;
						if (use_gaussian) {
							float dist = std::hypot(di_factor * i + sub_halfwidth, di_factor * j + sub_halfwidth) / max_dist;
							weight     = std::exp((double)(-(dist * dist) * rsigma_squared));
						} else if (use_tukey) {
							if (S / 2 - std::abs(S / 2 - i) <= sigma * S / 2)
								weight *= 0.5f * (1.f - cosf(2.f * M_PI * (S / 2 - std::abs(S / 2 - i)) * rtukey));
							if (S / 2 - std::abs(S / 2 - j) <= sigma * S / 2)
								weight *= 0.5f * (1.f - cosf(2.f * M_PI * (S / 2 - std::abs(S / 2 - j)) * rtukey));
						} else if (use_square) {
							float inv_dist =
							    1 - std::hypot(di_factor * i + sub_halfwidth, di_factor * j + sub_halfwidth) / max_dist;
							weight = inv_dist * inv_dist;
						}
						avg_pressure += weight * raw;
					}
				}
			}
			out[x + cy * y] = avg_pressure;
		}
	}
	if (ext)
		g_ext.destroy(ext);
}

void set_external_caster(const ExternalCaster &c) { g_ext = c; }

// The oracle's own BLAS/TLAS/Moeller-Trumbore restatement on caller-supplied triangle soups and rays: the entry the
// tests use to hold it against the reference's compiled ray caster (oracle/_ref) and the vectors made with it
// (tests/golden/ref_bvh_*.npz).  Same argument layout as ref_tlas_create / ref_tlas_cast (oracle/ref_shim/ref_capi.cpp).
void cast_rays(int n_surf, const int *n_tri, const double *verts, int n_rays, const float *O, const float *D,
               float *tuv, uint32_t *id)
{
	std::vector<Surface> surf(n_surf);
	size_t off = 0;
	for (int s = 0; s < n_surf; ++s) {
		Surface &sf = surf[s];
		sf.tri      = true;
		for (int t = 0; t < n_tri[s]; ++t, ++off) {
			sf.face_first.push_back((int)sf.face_idx.size());
			sf.face_n.push_back(3);
			for (int k = 0; k < 3; ++k) {
				sf.face_idx.push_back((int)sf.v.size());
				sf.v.push_back({ verts[9 * off + 3 * k], verts[9 * off + 3 * k + 1], verts[9 * off + 3 * k + 2] });
			}
		}
	}
	std::vector<Blas> blas(n_surf);
	for (int s = 0; s < n_surf; ++s)
		blas[s].build(surf[s]);
	Tlas tlas;
	tlas.build(blas);
	for (int i = 0; i < n_rays; ++i) {
		Ray ray;
		ray.O = { O[3 * i], O[3 * i + 1], O[3 * i + 2] };
		ray.D = { D[3 * i], D[3 * i + 1], D[3 * i + 2] };
		ray.t = 1e30f, ray.u = ray.v = 0, ray.hit = 0;
		tlas.intersect(ray);
		tuv[3 * i] = ray.t, tuv[3 * i + 1] = ray.u, tuv[3 * i + 2] = ray.v;
		id[i] = ray.hit;
	}
}

// single primitives (bvh.cpp:49-74, bvh.h:157-176), for the known-answer comparison with oracle/_ref
void intersect_triangle_one(const float *O, const float *D, const float *v0, const float *v1, const float *v2, float t_in,
                            float *tuv_out, int *hit_out)
{
	Ray ray;
	ray.O = { O[0], O[1], O[2] }, ray.D = { D[0], D[1], D[2] };
	ray.t = t_in, ray.u = ray.v = 0, ray.hit = 0xffffffffu;
	Tri tri{ { v0[0], v0[1], v0[2] }, { v1[0], v1[1], v1[2] }, { v2[0], v2[1], v2[2] }, { 0, 0, 0 } };
	intersect_triangle(ray, tri, 7u);
	tuv_out[0] = ray.t, tuv_out[1] = ray.u, tuv_out[2] = ray.v;
	*hit_out   = ray.hit == 7u;
}
float intersect_aabb_one(const float *O, const float *D, float t_in, const float *bmin, const float *bmax)
{
	Ray ray;
	ray.O = { O[0], O[1], O[2] }, ray.D = { D[0], D[1], D[2] };
	ray.rD = { 1.0f / D[0], 1.0f / D[1], 1.0f / D[2] };
	ray.t  = t_in;
	return intersect_aabb(ray, { bmin[0], bmin[1], bmin[2] }, { bmax[0], bmax[1], bmax[2] });
}

// curved_sensor.cpp:325-368: distance matrix taxel x sample, assignment of every sample within include_margin of a
// taxel (and, when the taxel has a normal, within 45 degrees of it), weight (include_margin - distance)^2, "close"
// samples kept in the order they are first used.  dist(i,j) is written as the reference's Eigen expression
// (-2 t.s + |t|^2) + |s|^2; the summation order inside Eigen's product is not observable.
void curved_sensor_load(CurvedSensor &cs, const double *sample_pos, const double *sample_nrm, int m)
{
	const int n = (int)cs.taxel_pos.size();
	cs.surface_idx.assign(n, {});
	cs.surface_weight.assign(n, {});
	cs.surf_pos.clear(), cs.surf_nrm.clear(), cs.surf_src.clear();
	const double margin_sq = cs.include_margin * cs.include_margin;
	for (int j = 0; j < m; ++j) {
		V3 s{ sample_pos[3 * j], sample_pos[3 * j + 1], sample_pos[3 * j + 2] };
		V3 sn{ sample_nrm[3 * j], sample_nrm[3 * j + 1], sample_nrm[3 * j + 2] };
		bool added = false;
		for (int i = 0; i < n; ++i) {
			const V3 &t = cs.taxel_pos[i], &tn = cs.taxel_nrm[i];
			double dist = (-2 * (t[0] * s[0] + t[1] * s[1] + t[2] * s[2]) + (t[0] * t[0] + t[1] * t[1] + t[2] * t[2])) +
			              (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
			double nsq = tn[0] * tn[0] + tn[1] * tn[1] + tn[2] * tn[2];
			if (dist < margin_sq &&
			    (nsq == 0 || std::acos(tn[0] * sn[0] + tn[1] * sn[1] + tn[2] * sn[2]) < 45 * M_PI / 180.)) {
				if (!added) {
					cs.surf_pos.push_back(s), cs.surf_nrm.push_back(sn), cs.surf_src.push_back(j);
					added = true;
				}
				cs.surface_idx[i].push_back((int)cs.surf_pos.size() - 1);
				cs.surface_weight[i].push_back(std::pow(std::max(0.0, cs.include_margin - std::sqrt(dist)), 2));
			}
		}
	}
}

// curved_sensor.cpp:388-481 internal_update: one ray per (taxel, assigned sample) from the sample point along the
// inward normal, nearest hit with 0 < t < include_margin, pressure = sum of weight * e_MN(hit).
void curved_sensor_values(const Scene &sc, const StepState &st, int sensor, float *out, int caster)
{
	const bool use_bvh = caster == 1, use_ext = caster == 2;
	const CurvedSensor &cs = sc.curved[sensor];
	const int id = cs.geom, n = (int)cs.taxel_pos.size();
	double rot[9];
	std::memcpy(rot, &st.xmat[9 * id], sizeof(rot));
	const double *xp = &st.xpos[3 * id];
	std::fill(out, out + n, 0.0f);
	std::vector<Blas> blas;
	for (const PairOut &po : st.out)
		if (po.has_surface && (po.gM == id || po.gN == id) && po.s->tri)
			blas.emplace_back();
	{
		int b = 0;
		for (const PairOut &po : st.out)
			if (po.has_surface && (po.gM == id || po.gN == id) && po.s->tri)
				blas[b++].build(*po.s);
	}
	if (blas.empty())
		return;
	Tlas tlas;
	tlas.build(blas);
	void *ext = use_ext ? make_external_tlas(blas) : nullptr;
	for (int i = 0; i < n; ++i) {
		double pressure = 0;
		for (size_t j0 = 0; j0 < cs.surface_idx[i].size(); ++j0) {
			int j = cs.surface_idx[i][j0];
			const V3 &p = cs.surf_pos[j], &nn = cs.surf_nrm[j];
			// M * (p, 1) and M * (n, 0) with M = [R | x] (:391-399)
			double w[3], wn[3];
			for (int r = 0; r < 3; ++r) {
				w[r]  = rot[3 * r] * p[0] + rot[3 * r + 1] * p[1] + rot[3 * r + 2] * p[2] + xp[r];
				wn[r] = rot[3 * r] * nn[0] + rot[3 * r + 1] * nn[1] + rot[3 * r + 2] * nn[2];
			}
			F3 normal{ (float)wn[0], (float)wn[1], (float)wn[2] };
			F3 surface_point{ (float)w[0], (float)w[1], (float)w[2] };
			Ray ray;
			ray.O   = surface_point + normal * (float)1e-8;
			ray.D   = -normal;
			ray.t   = 1e30f;
			ray.u = ray.v = 0;
			ray.hit = 0;
			if (use_ext) {
				external_cast(ext, ray);
			} else if (use_bvh) {
				tlas.intersect(ray);
			} else {
				ray.rD = { 1.0f / ray.D.x, 1.0f / ray.D.y, 1.0f / ray.D.z };
				for (unsigned b = 0; b < blas.size(); ++b)
					for (unsigned t = 0; t < blas[b].tri.size(); ++t)
						intersect_triangle(ray, blas[b].tri[t], (b << 20) + t);
			}
			if (ray.t < cs.include_margin && ray.t > 0.0f) {
				double bary[3]   = { (double)(1 - ray.u - ray.v), (double)ray.u, (double)ray.v };
				unsigned tri_idx = ray.hit & 0xFFFFF, blas_idx = ray.hit >> 20;
				const Surface &s = *blas[blas_idx].surface;
				const int *f     = &s.face_idx[s.face_first[tri_idx]];
				double raw       = bary[0] * s.e[f[0]];
				raw += bary[1] * s.e[f[1]];
				raw += bary[2] * s.e[f[2]];
				pressure += cs.surface_weight[i][j0] * raw;
			}
		}
		out[i] = (float)pressure;
	}
	if (ext)
		g_ext.destroy(ext);
}

// taxel_sensor.cpp:158-478 internal_update, DEFAULT sampling.  Reproduced as written, including quirk Q12: the
// method switch has no breaks (weighted -> mean -> squared: all three end with the squared result), closest keeps
// the pressure only when visualize is on, st2 is computed from the same edge as st1, and taxels without a sample in
// range keep the value of the previous update (only "no sample at all" writes zeros).
void taxel_sensor_values(const Scene &sc, const StepState &st, int sensor, float *values)
{
	const TaxelSensor &ts = sc.taxel[sensor];
	const int id = ts.geom, n = (int)ts.taxels.size();
	const double margin = ts.include_margin, margin_sq = margin * margin, res = ts.sample_resolution;
	std::vector<V3> spoints;
	std::vector<double> spress; // s->tri_e_MN().Evaluate(t, bary) of each sample
	// AREA_IMPORTANCE (taxel_sensor.cpp:211-254): ONE std::default_random_engine per update (default seed, so every
	// update draws the same stream), one stratum of sample_resolution * total_area per sample along the cumulative
	// triangle area, uniform barycentric point inside the triangle that owns the stratum's start.  Which random numbers
	// a triangle gets depends on the triangle ORDER, and Drake's is unobservable; the order here is canonical: pairs in
	// pair order, polygons by (elemM, elemN), fan triangles in fan order, and area(t) = |(b - a) x (c - a)| / 2 of the
	// WORLD vertices (Drake computes it in the builder frame: ulps), total_area their sum in that order.
	std::default_random_engine generator;
	std::uniform_real_distribution<double> distribution(0.0, 1.0);
	for (const PairOut &po : st.out) {
		if (ts.sample_method != 1)
			break;
		if (!(po.has_surface && (po.gM == id || po.gN == id) && po.s->tri))
			continue;
		const Surface &s = *po.s;
		std::vector<std::array<int, 4>> polys; // (elemM, elemN, first fan triangle, fan size)
		for (const Emitted &em : s.emitted)
			polys.push_back({ em.elemM, em.elemN, em.first_face, em.n_faces });
		std::sort(polys.begin(), polys.end());
		std::vector<int> order;
		for (const auto &pl : polys)
			for (int i = 0; i < pl[3]; ++i)
				order.push_back(pl[2] + i);
		auto tri_area = [&](int t) {
			const int *f = &s.face_idx[s.face_first[t]];
			return 0.5 * norm(cross(s.v[f[1]] - s.v[f[0]], s.v[f[2]] - s.v[f[0]]));
		};
		double total = 0;
		for (int t : order)
			total += tri_area(t);
		const double area_resolution = res * total;
		if (!(area_resolution > 0))
			continue;
		double area = 0, at = 0;
		for (int t : order) {
			const int *f = &s.face_idx[s.face_first[t]];
			at += tri_area(t);
			while (area < at) {
				area += area_resolution;
				double u0 = distribution(generator), u1 = distribution(generator);
				double a = 1.0 - std::sqrt(u0), b = (1.0 - a) * u1;
				double bary[3] = { a, (1 - a) * (1 - b), (1 - a) * b };
				V3 p;
				for (int k = 0; k < 3; ++k)
					p.at(k) = bary[0] * s.v[f[0]][k] + bary[1] * s.v[f[1]][k] + bary[2] * s.v[f[2]][k];
				spoints.push_back(p);
				spress.push_back(bary[0] * s.e[f[0]] + bary[1] * s.e[f[1]] + bary[2] * s.e[f[2]]);
			}
		}
	}
	for (const PairOut &po : st.out) {
		if (ts.sample_method != 0)
			break;
		if (!(po.has_surface && (po.gM == id || po.gN == id) && po.s->tri))
			continue;
		const Surface &s = *po.s;
		for (int t = 0; t < s.num_faces(); ++t) {
			const int *f = &s.face_idx[s.face_first[t]];
			const V3 &v0 = s.v[f[0]], &v1 = s.v[f[1]], &v2 = s.v[f[2]];
			auto norm = [](const V3 &a, const V3 &b) {
				double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
				return std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
			};
			int st0 = (int)(norm(v1, v0) / res) + 1;
			int st1 = (int)(norm(v2, v0) / res) + 1;
			int st2 = (int)(norm(v2, v0) / res) + 1;
			int stm = std::max(st0, st1);
			for (double a = 0; a <= 1; a += 1. / stm)
				for (double b = 0; b <= 1; b += 1. / st2) {
					double bary[3] = { a, (1 - a) * (1 - b), (1 - a) * b };
					V3 p;
					for (int k = 0; k < 3; ++k)
						p.at(k) = bary[0] * v0[k] + bary[1] * v1[k] + bary[2] * v2[k];
					spoints.push_back(p);
					spress.push_back(bary[0] * s.e[f[0]] + bary[1] * s.e[f[1]] + bary[2] * s.e[f[2]]);
				}
		}
	}
	const int m = (int)spoints.size();
	if (m == 0) {
		std::fill(values, values + n, 0.0f);
		return;
	}
	const double *R = &st.xmat[9 * id], *xp = &st.xpos[3 * id];
	for (int i = 0; i < n; ++i) {
		V3 tw;
		for (int r = 0; r < 3; ++r)
			tw.at(r) = R[3 * r] * ts.taxels[i][0] + R[3 * r + 1] * ts.taxels[i][1] + R[3 * r + 2] * ts.taxels[i][2] + xp[r];
		auto dist = [&](int j) { // (-2 t.s + |t|^2) + |s|^2, :286-289
			const V3 &sp = spoints[j];
			return (-2 * (tw[0] * sp[0] + tw[1] * sp[1] + tw[2] * sp[2]) + (tw[0] * tw[0] + tw[1] * tw[1] + tw[2] * tw[2])) +
			       (sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2]);
		};
		if (ts.method == 0) { // closest
			int jmin = 0;
			double dmin = dist(0);
			for (int j = 1; j < m; ++j) {
				double dj = dist(j);
				if (dj < dmin)
					dmin = dj, jmin = j;
			}
			if (dmin < margin_sq) {
				double pressure = spress[jmin];
				values[i]       = (float)pressure;
				if (!(ts.visualize && std::abs(pressure) > 1e-6))
					values[i] = 0;
			}
			continue;
		}
		// weighted / mean fall through to squared: only its assignment survives
		double ws = 0, pressure = 0;
		for (int j = 0; j < m; ++j) {
			double dj = dist(j);
			if (dj < margin_sq) {
				double w = std::pow(std::max(0.0, margin - std::sqrt(dj)), 2);
				pressure += w * std::abs(spress[j]);
				ws += 1;
			}
		}
		if (ws > 0) {
			pressure *= res;
			values[i] = (float)pressure;
		}
	}
}

} // namespace orc
