// Host-side mesh generation for the B200 hydroelastic engine (see mesh_host.h).
// Behaviour follows mujoco_contact_surfaces_plugin.cpp:633-808 (which shapes exist, how MuJoCo
// sizes map to shapes, soft vs rigid) and the Drake v1.8.0 mesh rules restated in SURVEY.md App. A.1/A.2.
#include "mesh_host.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <unordered_map>

namespace hcs {
namespace {

enum { GEOM_PLANE = 0, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH };

struct P3 {
	double x, y, z;
};
inline P3 sub(P3 a, P3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline P3 crs(P3 a, P3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline double dt(P3 a, P3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }

struct Builder {
	std::vector<P3> v;
	std::vector<int32_t> e;
	int add(P3 p)
	{
		v.push_back(p);
		return (int)v.size() - 1;
	}
	double six_volume(int a, int b, int c, int d) const { return dt(crs(sub(v[b], v[a]), sub(v[c], v[a])), sub(v[d], v[a])); }
	void tet(int a, int b, int c, int d)
	{
		e.push_back(a), e.push_back(b), e.push_back(c), e.push_back(d);
	}
	void flush(HostMesh &m) const
	{
		m.verts.clear();
		for (const P3 &p : v)
			m.verts.push_back(p.x), m.verts.push_back(p.y), m.verts.push_back(p.z);
		m.elems = e;
	}
};

// refinement level from the resolution hint: the level-L refined octahedron has 4*2^L equator
// edges; the chord e = min(hint, 2r) subtends 2*asin(e/2r).
int sphere_level(double r, double hint)
{
	double e  = std::min(hint, 2.0 * r);
	int level = (int)std::ceil(std::log2(M_PI / std::asin(e / (2.0 * r)))) - 2;
	return level < 0 ? 0 : level;
}

// unit sphere, single interior vertex (index 0); boundary triangles refined `level` times.
void unit_sphere(int level, Builder &b, std::vector<int32_t> &boundary)
{
	b.v      = { { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 }, { -1, 0, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } };
	boundary = { 1, 2, 5, 2, 3, 5, 3, 4, 5, 4, 1, 5, 2, 1, 6, 3, 2, 6, 4, 3, 6, 1, 4, 6 };
	for (int l = 0; l < level; ++l) {
		std::unordered_map<uint64_t, int> mid;
		mid.reserve(boundary.size());
		auto M = [&](int p, int q) {
			uint64_t key = ((uint64_t)std::min(p, q) << 32) | (uint32_t)std::max(p, q);
			auto it      = mid.find(key);
			if (it != mid.end())
				return it->second;
			const P3 &A = b.v[std::min(p, q)], &B = b.v[std::max(p, q)];
			P3 m     = { (A.x + B.x) * 0.5, (A.y + B.y) * 0.5, (A.z + B.z) * 0.5 };
			double z = dt(m, m);
			if (z > 0) {
				double s = std::sqrt(z);
				m        = { m.x / s, m.y / s, m.z / s };
			}
			int id   = b.add(m);
			mid[key] = id;
			return id;
		};
		std::vector<int32_t> next;
		next.reserve(boundary.size() * 4);
		for (size_t t = 0; t < boundary.size(); t += 3) {
			int a = boundary[t], bb = boundary[t + 1], c = boundary[t + 2];
			int ab = M(a, bb), bc = M(bb, c), ca = M(c, a);
			int32_t four[12] = { a, ab, ca, ab, bb, bc, ca, bc, c, ab, bc, ca };
			next.insert(next.end(), four, four + 12);
		}
		boundary.swap(next);
	}
}

void sphere_like(double a, double bsz, double c, int level, Builder &b)
{
	std::vector<int32_t> boundary;
	unit_sphere(level, b, boundary);
	for (P3 &p : b.v)
		p = { p.x * a, p.y * bsz, p.z * c };
	for (size_t t = 0; t < boundary.size(); t += 3)
		b.tet(0, boundary[t], boundary[t + 1], boundary[t + 2]);
}

void box_grid(double sx, double sy, double sz, double hint, Builder &b)
{
	const double s[3] = { sx, sy, sz };
	int n[3];
	for (int a = 0; a < 3; ++a)
		n[a] = std::max(1, (int)std::ceil(s[a] / hint));
	auto coord = [&](int a, int i) { return i == n[a] ? s[a] / 2 : -(s[a] / 2) + i * (s[a] / n[a]); };
	for (int i = 0; i <= n[0]; ++i)
		for (int j = 0; j <= n[1]; ++j)
			for (int k = 0; k <= n[2]; ++k)
				b.add({ coord(0, i), coord(1, j), coord(2, k) });
	auto id = [&](int i, int j, int k) { return (i * (n[1] + 1) + j) * (n[2] + 1) + k; };
	// hexagon of cell corners around the main diagonal 000-111
	const int ring[7][3] = { { 0, 0, 1 }, { 0, 1, 1 }, { 0, 1, 0 }, { 1, 1, 0 }, { 1, 0, 0 }, { 1, 0, 1 }, { 0, 0, 1 } };
	for (int i = 0; i < n[0]; ++i)
		for (int j = 0; j < n[1]; ++j)
			for (int k = 0; k < n[2]; ++k)
				for (int q = 0; q < 6; ++q)
					b.tet(id(i + ring[q][0], j + ring[q][1], k + ring[q][2]),
					      id(i + ring[q + 1][0], j + ring[q + 1][1], k + ring[q + 1][2]), id(i + 1, j + 1, k + 1), id(i, j, k));
}

// (possibly degenerate) hexahedral cell -> tets: fan from the smallest global id over the
// triangles of the faces not containing it; quads are cut through their own smallest global id.
void cell_to_tets(Builder &b, const int H[8], double vol_eps)
{
	static const int quad[6][4] = { { 0, 1, 3, 2 }, { 4, 5, 7, 6 }, { 0, 1, 5, 4 },
		                             { 2, 3, 7, 6 }, { 0, 2, 6, 4 }, { 1, 3, 7, 5 } };
	int apex = H[0];
	for (int q = 1; q < 8; ++q)
		apex = std::min(apex, H[q]);
	for (int f = 0; f < 6; ++f) {
		int g[4] = { H[quad[f][0]], H[quad[f][1]], H[quad[f][2]], H[quad[f][3]] };
		if (g[0] == apex || g[1] == apex || g[2] == apex || g[3] == apex)
			continue;
		int lo = 0;
		for (int q = 1; q < 4; ++q)
			if (g[q] < g[lo])
				lo = q;
		int r[4] = { g[lo], g[(lo + 1) % 4], g[(lo + 2) % 4], g[(lo + 3) % 4] };
		for (int half = 0; half < 2; ++half) {
			int t0 = r[0], t1 = r[1 + half], t2 = r[2 + half];
			if (t0 == t1 || t1 == t2 || t0 == t2)
				continue;
			double sv = b.six_volume(apex, t0, t1, t2);
			if (std::fabs(sv) <= vol_eps)
				continue;
			if (sv > 0)
				b.tet(apex, t0, t1, t2);
			else
				b.tet(apex, t0, t2, t1);
		}
	}
}

void box_medial(double sx, double sy, double sz, Builder &b)
{
	const double h[3] = { sx / 2, sy / 2, sz / 2 };
	const double m    = std::min(h[0], std::min(h[1], h[2]));
	int corner[8], medial[8];
	for (int c = 0; c < 8; ++c)
		corner[c] = b.add({ (c & 4) ? h[0] : -h[0], (c & 2) ? h[1] : -h[1], (c & 1) ? h[2] : -h[2] });
	for (int c = 0; c < 8; ++c) {
		double d[3] = { h[0] - m, h[1] - m, h[2] - m };
		P3 p        = { d[0] == 0 ? 0.0 : ((c & 4) ? d[0] : -d[0]), d[1] == 0 ? 0.0 : ((c & 2) ? d[1] : -d[1]),
			            d[2] == 0 ? 0.0 : ((c & 1) ? d[2] : -d[2]) };
		int hit     = -1;
		for (int q = 8; q < (int)b.v.size(); ++q)
			if (b.v[q].x == p.x && b.v[q].y == p.y && b.v[q].z == p.z)
				hit = q;
		medial[c] = hit >= 0 ? hit : b.add(p);
	}
	const double eps = 1e-13 * (sx * sy * sz);
	const int bit[3] = { 4, 2, 1 };
	for (int axis = 0; axis < 3; ++axis)
		for (int side = 0; side < 2; ++side) {
			int H[8];
			for (int o = 0; o < 2; ++o)
				for (int u = 0; u < 2; ++u)
					for (int w = 0; w < 2; ++w) {
						int c = (side ? bit[axis] : 0) | (u ? bit[(axis + 1) % 3] : 0) | (w ? bit[(axis + 2) % 3] : 0);
						H[4 * o + 2 * u + w] = o ? medial[c] : corner[c];
					}
			cell_to_tets(b, H, eps);
		}
}

bool cylinder_medial(double r, double length, double hint, Builder &b, std::string &err)
{
	if (!(hint > 0)) {
		err = "soft cylinder needs resolutionHint > 0 (the reference leaves the mesh uninitialised, plugin.cpp:698-701)";
		return false;
	}
	const int n    = std::max(3, (int)std::ceil(2.0 * M_PI * r / hint));
	const double h = length / 2;
	std::vector<int> B(n), T(n);
	for (int i = 0; i < n; ++i) {
		double th = 2.0 * M_PI * i / n;
		B[i]      = b.add({ r * std::cos(th), r * std::sin(th), -h });
	}
	for (int i = 0; i < n; ++i)
		T[i] = b.add({ b.v[B[i]].x, b.v[B[i]].y, h });
	int Cb = b.add({ 0, 0, -h }), Ct = b.add({ 0, 0, h });
	const double eps = 1e-13 * (r * r * length);
	if (h >= r) {
		int M0 = b.add({ 0, 0, h == r ? 0.0 : -(h - r) });
		int M1 = h > r ? b.add({ 0, 0, h - r }) : M0;
		for (int i = 0; i < n; ++i) {
			int j     = (i + 1) % n;
			int Hb[8] = { Cb, Cb, B[i], B[j], M0, M0, M0, M0 };
			int Ht[8] = { Ct, Ct, T[i], T[j], M1, M1, M1, M1 };
			int Hs[8] = { B[i], T[i], B[j], T[j], M0, M1, M0, M1 };
			cell_to_tets(b, Hb, eps);
			cell_to_tets(b, Ht, eps);
			cell_to_tets(b, Hs, eps);
		}
	} else {
		const double rm = r - h;
		int Mc          = b.add({ 0, 0, 0 });
		std::vector<int> M(n);
		for (int i = 0; i < n; ++i) {
			double th = 2.0 * M_PI * i / n;
			M[i]      = b.add({ rm * std::cos(th), rm * std::sin(th), 0 });
		}
		for (int i = 0; i < n; ++i) {
			int j     = (i + 1) % n;
			int Hb[8] = { Cb, Cb, B[i], B[j], Mc, Mc, M[i], M[j] };
			int Ht[8] = { Ct, Ct, T[i], T[j], Mc, Mc, M[i], M[j] };
			int Hs[8] = { B[i], T[i], B[j], T[j], M[i], M[i], M[j], M[j] };
			cell_to_tets(b, Hb, eps);
			cell_to_tets(b, Ht, eps);
			cell_to_tets(b, Hs, eps);
		}
	}
	return true;
}

// boundary of a tet mesh as an outward-wound triangle surface; vertices renumbered ascending
void boundary_surface(const Builder &vol, HostMesh &out)
{
	static const int face[4][3] = { { 1, 2, 3 }, { 0, 3, 2 }, { 0, 1, 3 }, { 0, 2, 1 } };
	struct Key {
		int a, b, c;
		bool operator<(const Key &o) const { return a != o.a ? a < o.a : (b != o.b ? b < o.b : c < o.c); }
	};
	auto key_of = [](int a, int b, int c) {
		int s[3] = { a, b, c };
		std::sort(s, s + 3);
		return Key{ s[0], s[1], s[2] };
	};
	std::map<Key, int> uses;
	size_t nt = vol.e.size() / 4;
	for (size_t t = 0; t < nt; ++t)
		for (auto &f : face)
			uses[key_of(vol.e[4 * t + f[0]], vol.e[4 * t + f[1]], vol.e[4 * t + f[2]])]++;
	std::vector<int32_t> tris;
	std::vector<int> newid(vol.v.size(), -1);
	for (size_t t = 0; t < nt; ++t)
		for (auto &f : face) {
			int a = vol.e[4 * t + f[0]], b = vol.e[4 * t + f[1]], c = vol.e[4 * t + f[2]];
			if (uses[key_of(a, b, c)] == 1) {
				tris.push_back(a), tris.push_back(b), tris.push_back(c);
				newid[a] = newid[b] = newid[c] = 0;
			}
		}
	out.verts.clear();
	int next = 0;
	for (size_t i = 0; i < vol.v.size(); ++i)
		if (newid[i] == 0) {
			newid[i] = next++;
			out.verts.push_back(vol.v[i].x), out.verts.push_back(vol.v[i].y), out.verts.push_back(vol.v[i].z);
		}
	out.elems.resize(tris.size());
	for (size_t i = 0; i < tris.size(); ++i)
		out.elems[i] = newid[tris[i]];
}

inline double snap_extent(double x) { return std::fabs(x) < 1e-14 ? 0.0 : x; }

} // namespace

bool build_geom_mesh(int mj_type, const double size[3], const float *mesh_vert, int n_vert, const int32_t *mesh_face,
                     int n_face, const double props[5], HostMesh &out, std::string &err)
{
	const double E = props[0], hint = props[2];
	const bool soft = E > 0;
	out             = HostMesh();
	out.soft        = soft;
	Builder vol;
	switch (mj_type) {
		case GEOM_PLANE:
			if (soft) {
				err = "soft plane collision not implemented (plugin.cpp:635-636)";
				return false;
			}
			out.plane = true;
			return true;
		case GEOM_SPHERE:
		case GEOM_ELLIPSOID: {
			if (!(hint > 0)) {
				err = "sphere/ellipsoid need resolutionHint > 0 (Drake DRAKE_DEMAND)";
				return false;
			}
			double a = size[0], bsz = mj_type == GEOM_SPHERE ? size[0] : size[1], c = mj_type == GEOM_SPHERE ? size[0] : size[2];
			int level;
			if (mj_type == GEOM_SPHERE)
				level = sphere_level(a, hint);
			else
				level = sphere_level(1.0, hint / std::max(a, std::max(bsz, c)));
			sphere_like(a, bsz, c, level, vol);
			if (soft) {
				for (const P3 &p : vol.v) {
					P3 q = mj_type == GEOM_SPHERE ? p : P3{ p.x / a, p.y / bsz, p.z / c };
					double rad = std::sqrt(dt(q, q));
					double ext = mj_type == GEOM_SPHERE ? 1.0 - rad / a : 1.0 - rad;
					out.pressure.push_back(E * snap_extent(ext));
				}
			}
			break;
		}
		case GEOM_CYLINDER: {
			double r = size[0], length = 2 * size[1];
			if (!cylinder_medial(r, length, hint, vol, err))
				return false;
			if (soft) {
				double h = length / 2, m = std::min(r, h);
				for (const P3 &p : vol.v) {
					double rho = std::sqrt(p.x * p.x + p.y * p.y);
					out.pressure.push_back(E * snap_extent(std::min(r - rho, h - std::fabs(p.z)) / m));
				}
			}
			break;
		}
		case GEOM_BOX: {
			double sx = 2 * size[0], sy = 2 * size[1], sz = 2 * size[2];
			if (soft && !(hint > 0))
				box_medial(sx, sy, sz, vol);
			else if (hint > 0)
				box_grid(sx, sy, sz, hint, vol);
			else {
				err = "rigid box needs resolutionHint > 0 (MakeBoxSurfaceMesh)";
				return false;
			}
			if (soft) {
				double h[3] = { sx / 2, sy / 2, sz / 2 };
				double m    = std::min(h[0], std::min(h[1], h[2]));
				for (const P3 &p : vol.v) {
					double d = std::min(h[0] - std::fabs(p.x), std::min(h[1] - std::fabs(p.y), h[2] - std::fabs(p.z)));
					out.pressure.push_back(E * snap_extent(d / m));
				}
			}
			break;
		}
		case GEOM_MESH: {
			if (n_vert <= 0 || n_face <= 0 || !mesh_vert || !mesh_face) {
				err = "Could not load mesh! It was not properly defined in mujoco. (plugin.cpp:804-805)";
				return false;
			}
			for (int i = 0; i < n_vert; ++i)
				vol.add({ (double)mesh_vert[3 * i], (double)mesh_vert[3 * i + 1], (double)mesh_vert[3 * i + 2] });
			if (!soft) {
				vol.flush(out);
				out.elems.assign(mesh_face, mesh_face + 3 * (size_t)n_face);
				return true;
			}
			// centroid of the enclosed volume (plugin.cpp:161-187) and the centroid fan (:767-787)
			double six_total = 0;
			P3 acc{ 0, 0, 0 };
			for (int f = 0; f < n_face; ++f) {
				const P3 &p = vol.v[mesh_face[3 * f]], &q = vol.v[mesh_face[3 * f + 1]], &r = vol.v[mesh_face[3 * f + 2]];
				double sv   = dt(crs(p, q), r);
				six_total += sv;
				P3 s = { (p.x + q.x) + r.x, (p.y + q.y) + r.y, (p.z + q.z) + r.z };
				acc  = { acc.x + sv * s.x, acc.y + sv * s.y, acc.z + sv * s.z };
			}
			double den = 4 * six_total;
			int ci     = vol.add({ acc.x / den, acc.y / den, acc.z / den });
			for (int f = 0; f < n_face; ++f)
				vol.tet(ci, mesh_face[3 * f], mesh_face[3 * f + 1], mesh_face[3 * f + 2]);
			out.pressure.assign(vol.v.size(), 0.0);
			out.pressure.back() = E; // MakeConvexPressureField
			break;
		}
		default:
			err = mj_type == GEOM_HFIELD ? "hfield collision not implemented yet (plugin.cpp:642-644)" :
			                                "capsule collision not implemented yet (plugin.cpp:645-647)";
			return false;
	}
	if (soft)
		vol.flush(out);
	else
		boundary_surface(vol, out);
	return true;
}

int sphere_like_level(int mj_type, const double size[3], double hint)
{
	if (mj_type == GEOM_SPHERE)
		return sphere_level(size[0], hint);
	return sphere_level(1.0, hint / std::max(size[0], std::max(size[1], size[2])));
}

void unit_sphere_vertices(int level, std::vector<double> &verts)
{
	Builder b;
	std::vector<int32_t> boundary;
	unit_sphere(level, b, boundary);
	verts.clear();
	for (const P3 &p : b.v)
		verts.push_back(p.x), verts.push_back(p.y), verts.push_back(p.z);
}

} // namespace hcs
