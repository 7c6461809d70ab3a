// CPU ORACLE (test infrastructure, parity unpinned) — mesh and pressure-field construction.
//
// Restates what mujoco_contact_surfaces_plugin.cpp:613-812 obtains from Drake v1.8.0
// geometry/proximity/{make_sphere_mesh.h, make_ellipsoid_mesh.cc, make_box_mesh.cc,
// make_cylinder_mesh.cc, make_*_field.h, make_convex_field.h, volume_to_surface_mesh.cc,
// volume_mesh.h (CalcGradBarycentric), mesh_field_linear.h}.  Element/vertex ORDER is defined
// here (DESIGN.md "Mesh specification"); Drake's own enumeration order is not observable through
// the reference and does not change the pressure field or any surface integral.
#include "oracle.hpp"

#include <algorithm>
#include <map>
#include <stdexcept>

namespace orc {

// Drake make_sphere_mesh.h: edge chord e=min(hint,2r) subtends 2·asin(e/2r); the level-L octahedron
// refinement has 4·2^L equator edges  ⇒  L = ceil(log2(π / asin(e/2r))) − 2, clamped at 0.
int sphere_refinement_level(double r, double hint)
{
	if (!(hint > 0))
		throw std::runtime_error("sphere/ellipsoid resolution hint must be > 0");
	double e = std::min(hint, 2.0 * r);
	int L    = (int)std::ceil(std::log2(M_PI / std::asin(e / (2.0 * r)))) - 2;
	return std::max(0, L);
}

// kSingleInteriorVertex tessellation (plugin.cpp:653,677): vertex 0 is the centre, every tet is
// (centre, a, b, c) for an outward-wound boundary triangle (a,b,c).
VolumeMesh make_unit_sphere_volume(int level)
{
	std::vector<V3> v = { { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 }, { -1, 0, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } };
	std::vector<std::array<int, 3>> tri = { { 1, 2, 5 }, { 2, 3, 5 }, { 3, 4, 5 }, { 4, 1, 5 },
		                                     { 2, 1, 6 }, { 3, 2, 6 }, { 4, 3, 6 }, { 1, 4, 6 } };
	for (int l = 0; l < level; ++l) {
		std::map<std::pair<int, int>, int> mid;
		auto midpoint = [&](int a, int b) {
			std::pair<int, int> key(std::min(a, b), std::max(a, b));
			auto it = mid.find(key);
			if (it != mid.end())
				return it->second;
			V3 m = normalized((v[key.first] + v[key.second]) * 0.5);
			v.push_back(m);
			return mid[key] = (int)v.size() - 1;
		};
		std::vector<std::array<int, 3>> next;
		next.reserve(tri.size() * 4);
		for (auto &t : tri) {
			int ab = midpoint(t[0], t[1]), bc = midpoint(t[1], t[2]), ca = midpoint(t[2], t[0]);
			next.push_back({ t[0], ab, ca });
			next.push_back({ ab, t[1], bc });
			next.push_back({ ca, bc, t[2] });
			next.push_back({ ab, bc, ca });
		}
		tri.swap(next);
	}
	VolumeMesh vm;
	vm.v = v;
	for (auto &t : tri)
		vm.tets.push_back({ 0, t[0], t[1], t[2] });
	return vm;
}

VolumeMesh make_sphere_volume(double r, double hint)
{
	VolumeMesh vm = make_unit_sphere_volume(sphere_refinement_level(r, hint));
	for (auto &p : vm.v)
		p = p * r;
	return vm;
}

// Drake make_ellipsoid_mesh.cc: unit sphere at hint/max(a,b,c), scaled component-wise.
VolumeMesh make_ellipsoid_volume(double a, double b, double c, double hint)
{
	double r      = std::max(a, std::max(b, c));
	VolumeMesh vm = make_unit_sphere_volume(sphere_refinement_level(1.0, hint / r));
	for (auto &p : vm.v)
		p = { p.x * a, p.y * b, p.z * c };
	return vm;
}

static double six_vol(const std::vector<V3> &v, int a, int b, int c, int d)
{
	return dot(cross(v[b] - v[a], v[c] - v[a]), v[d] - v[a]);
}

// Drake make_box_mesh.cc MakeBoxVolumeMesh: n_i = ceil(size_i/hint) cells, six tets per cell around
// the cell's main diagonal.
VolumeMesh make_box_volume(double sx, double sy, double sz, double hint)
{
	if (!(hint > 0))
		throw std::runtime_error("box grid mesh needs hint > 0");
	int n[3]      = { std::max(1, (int)std::ceil(sx / hint)), std::max(1, (int)std::ceil(sy / hint)),
		              std::max(1, (int)std::ceil(sz / hint)) };
	double s[3]   = { sx, sy, sz };
	VolumeMesh vm;
	auto vid = [&](int i, int j, int k) { return (i * (n[1] + 1) + j) * (n[2] + 1) + k; };
	for (int i = 0; i <= n[0]; ++i)
		for (int j = 0; j <= n[1]; ++j)
			for (int k = 0; k <= n[2]; ++k) {
				int ijk[3] = { i, j, k };
				V3 p;
				for (int a = 0; a < 3; ++a)
					p.at(a) = ijk[a] == n[a] ? s[a] / 2 : -(s[a] / 2) + ijk[a] * (s[a] / n[a]);
				vm.v.push_back(p);
			}
	static const int cyc[7][3] = { { 0, 0, 1 }, { 0, 1, 1 }, { 0, 1, 0 }, { 1, 1, 0 }, { 1, 0, 0 }, { 1, 0, 1 }, { 0, 0, 1 } };
	for (int i = 0; i < n[0]; ++i)
		for (int j = 0; j < n[1]; ++j)
			for (int k = 0; k < n[2]; ++k) {
				int v000 = vid(i, j, k), v111 = vid(i + 1, j + 1, k + 1);
				for (int q = 0; q < 6; ++q) {
					int a = vid(i + cyc[q][0], j + cyc[q][1], k + cyc[q][2]);
					int b = vid(i + cyc[q + 1][0], j + cyc[q + 1][1], k + cyc[q + 1][2]);
					vm.tets.push_back({ a, b, v111, v000 });
				}
			}
	return vm;
}

// Split a (possibly degenerate) hexahedral cell H[4*o+2*u+v] into tets: fan from the cell's
// smallest global vertex id to the triangles of the faces that do not contain it, every quad face
// being split by the diagonal through ITS smallest global id.  The rule depends only on global ids,
// so neighbouring cells split their shared face identically (conforming mesh).  Tets with repeated
// vertices or zero volume are dropped; orientation is made positive.
static void split_cell(const std::vector<V3> &v, const int H[8], std::vector<std::array<int, 4>> &tets, double vol_eps)
{
	static const int faces[6][4] = { { 0, 1, 3, 2 }, { 4, 5, 7, 6 }, { 0, 1, 5, 4 },
		                              { 2, 3, 7, 6 }, { 0, 2, 6, 4 }, { 1, 3, 7, 5 } };
	int bs = 0;
	for (int b = 1; b < 8; ++b)
		if (H[b] < H[bs])
			bs = b;
	int apex = H[bs];
	for (auto &f : faces) {
		bool has = false;
		for (int q = 0; q < 4; ++q)
			has |= (H[f[q]] == apex);
		if (has)
			continue;
		int q0 = 0;
		for (int q = 1; q < 4; ++q)
			if (H[f[q]] < H[f[q0]])
				q0 = q;
		int g[4];
		for (int q = 0; q < 4; ++q)
			g[q] = H[f[(q0 + q) % 4]];
		int tri[2][3] = { { g[0], g[1], g[2] }, { g[0], g[2], g[3] } };
		for (auto &t : tri) {
			if (t[0] == t[1] || t[1] == t[2] || t[0] == t[2])
				continue;
			double sv = six_vol(v, apex, t[0], t[1], t[2]);
			if (std::fabs(sv) <= vol_eps)
				continue;
			if (sv > 0)
				tets.push_back({ apex, t[0], t[1], t[2] });
			else
				tets.push_back({ apex, t[0], t[2], t[1] });
		}
	}
}

// Drake MakeBoxVolumeMeshWithMa (plugin.cpp:726, hint==0): 8 corners + medial-axis vertices
// (±(hx−m), ±(hy−m), ±(hz−m)), m = min half size, coincident ones merged; one block per box face.
VolumeMesh make_box_volume_ma(double sx, double sy, double sz)
{
	double h[3] = { sx / 2, sy / 2, sz / 2 };
	double m    = std::min(h[0], std::min(h[1], h[2]));
	VolumeMesh vm;
	int corner[2][2][2], medial[2][2][2];
	for (int ix = 0; ix < 2; ++ix)
		for (int iy = 0; iy < 2; ++iy)
			for (int iz = 0; iz < 2; ++iz) {
				corner[ix][iy][iz] = (int)vm.v.size();
				vm.v.push_back({ ix ? h[0] : -h[0], iy ? h[1] : -h[1], iz ? h[2] : -h[2] });
			}
	for (int ix = 0; ix < 2; ++ix)
		for (int iy = 0; iy < 2; ++iy)
			for (int iz = 0; iz < 2; ++iz) {
				double d[3] = { h[0] - m, h[1] - m, h[2] - m };
				V3 p        = { d[0] == 0 ? 0.0 : (ix ? d[0] : -d[0]), d[1] == 0 ? 0.0 : (iy ? d[1] : -d[1]),
					            d[2] == 0 ? 0.0 : (iz ? d[2] : -d[2]) };
				int found   = -1;
				for (int q = 8; q < (int)vm.v.size(); ++q)
					if (vm.v[q].x == p.x && vm.v[q].y == p.y && vm.v[q].z == p.z)
						found = q;
				if (found < 0) {
					found = (int)vm.v.size();
					vm.v.push_back(p);
				}
				medial[ix][iy][iz] = found;
			}
	double vol_eps = 1e-13 * (sx * sy * sz);
	for (int axis = 0; axis < 3; ++axis)
		for (int side = 0; side < 2; ++side) {
			int H[8];
			for (int o = 0; o < 2; ++o)
				for (int u = 0; u < 2; ++u)
					for (int w = 0; w < 2; ++w) {
						int idx[3];
						idx[axis]           = side;
						idx[(axis + 1) % 3] = u;
						idx[(axis + 2) % 3] = w;
						H[4 * o + 2 * u + w] = o ? medial[idx[0]][idx[1]][idx[2]] : corner[idx[0]][idx[1]][idx[2]];
					}
			split_cell(vm.v, H, vm.tets, vol_eps);
		}
	return vm;
}

// Drake MakeCylinderVolumeMeshWithMa (plugin.cpp:700): n = max(3, ceil(2πr/hint)) rim vertices,
// medial axis = segment (half length > r), point (==), or disc of radius r − half length (<).
VolumeMesh make_cylinder_volume_ma(double r, double length, double hint)
{
	if (!(hint > 0))
		throw std::runtime_error("soft cylinder needs hint > 0 (reference leaves the mesh uninitialised, Q3)");
	int n    = std::max(3, (int)std::ceil(2.0 * M_PI * r / hint));
	double h = length / 2;
	VolumeMesh vm;
	std::vector<int> B(n), T(n);
	for (int i = 0; i < n; ++i) {
		double th = 2.0 * M_PI * i / n;
		B[i]      = (int)vm.v.size();
		vm.v.push_back({ r * std::cos(th), r * std::sin(th), -h });
	}
	for (int i = 0; i < n; ++i) {
		T[i] = (int)vm.v.size();
		vm.v.push_back({ vm.v[B[i]].x, vm.v[B[i]].y, h });
	}
	int Cb = (int)vm.v.size();
	vm.v.push_back({ 0, 0, -h });
	int Ct = (int)vm.v.size();
	vm.v.push_back({ 0, 0, h });
	double vol_eps = 1e-13 * (r * r * length);
	if (h >= r) { // long or medium: medial segment / point on the axis
		int M0 = (int)vm.v.size();
		vm.v.push_back({ 0, 0, h == r ? 0.0 : -(h - r) });
		int M1 = M0;
		if (h > r) {
			M1 = (int)vm.v.size();
			vm.v.push_back({ 0, 0, h - r });
		}
		for (int i = 0; i < n; ++i) {
			int j      = (i + 1) % n;
			int Hb[8]  = { Cb, Cb, B[i], B[j], M0, M0, M0, M0 };
			int Ht[8]  = { Ct, Ct, T[i], T[j], M1, M1, M1, M1 };
			int Hs[8]  = { B[i], T[i], B[j], T[j], M0, M1, M0, M1 };
			split_cell(vm.v, Hb, vm.tets, vol_eps);
			split_cell(vm.v, Ht, vm.tets, vol_eps);
			split_cell(vm.v, Hs, vm.tets, vol_eps);
		}
	} else { // short: medial disc of radius r-h in z=0
		double rm = r - h;
		int Mc    = (int)vm.v.size();
		vm.v.push_back({ 0, 0, 0 });
		std::vector<int> M(n);
		for (int i = 0; i < n; ++i) {
			double th = 2.0 * M_PI * i / n;
			M[i]      = (int)vm.v.size();
			vm.v.push_back({ rm * std::cos(th), rm * std::sin(th), 0 });
		}
		for (int i = 0; i < n; ++i) {
			int j     = (i + 1) % n;
			int Hb[8] = { Cb, Cb, B[i], B[j], Mc, Mc, M[i], M[j] };
			int Ht[8] = { Ct, Ct, T[i], T[j], Mc, Mc, M[i], M[j] };
			int Hs[8] = { B[i], T[i], B[j], T[j], M[i], M[i], M[j], M[j] };
			split_cell(vm.v, Hb, vm.tets, vol_eps);
			split_cell(vm.v, Ht, vm.tets, vol_eps);
			split_cell(vm.v, Hs, vm.tets, vol_eps);
		}
	}
	return vm;
}

// plugin.cpp:161-187 CalcCentroidOfEnclosedVolume + :767-787 centroid-fan tets {centroid,t0,t1,t2}
VolumeMesh make_convex_volume(const SurfaceMesh &sm)
{
	double six_total = 0;
	V3 c{ 0, 0, 0 };
	for (auto &t : sm.tris) {
		V3 p = sm.v[t[0]], q = sm.v[t[1]], r = sm.v[t[2]];
		double sv = dot(cross(p, q), r);
		six_total += sv;
		c = c + sv * ((p + q) + r);
	}
	c = c / (4 * six_total);
	VolumeMesh vm;
	vm.v = sm.v;
	vm.v.push_back(c);
	int ci = (int)vm.v.size() - 1;
	for (auto &t : sm.tris)
		vm.tets.push_back({ ci, t[0], t[1], t[2] });
	return vm;
}

void finish_surface(SurfaceMesh &sm)
{
	sm.normal.resize(sm.tris.size());
	sm.area.resize(sm.tris.size());
	for (size_t f = 0; f < sm.tris.size(); ++f) {
		V3 a = sm.v[sm.tris[f][0]], b = sm.v[sm.tris[f][1]], c = sm.v[sm.tris[f][2]];
		V3 cr     = cross(b - a, c - a);
		double nn = norm(cr);
		sm.area[f]   = 0.5 * nn;
		sm.normal[f] = nn != 0.0 ? cr / nn : cr;
	}
}

// Drake volume_to_surface_mesh.cc: faces that belong to exactly one tet, outward wound.
SurfaceMesh volume_to_surface(const VolumeMesh &vm)
{
	static const int F[4][3] = { { 1, 2, 3 }, { 0, 3, 2 }, { 0, 1, 3 }, { 0, 2, 1 } };
	std::map<std::array<int, 3>, int> count;
	for (auto &t : vm.tets)
		for (auto &f : F) {
			std::array<int, 3> k = { t[f[0]], t[f[1]], t[f[2]] };
			std::sort(k.begin(), k.end());
			count[k]++;
		}
	std::vector<std::array<int, 3>> tris;
	std::vector<char> used(vm.v.size(), 0);
	for (auto &t : vm.tets)
		for (auto &f : F) {
			std::array<int, 3> k = { t[f[0]], t[f[1]], t[f[2]] }, ks = k;
			std::sort(ks.begin(), ks.end());
			if (count[ks] == 1) {
				tris.push_back(k);
				used[k[0]] = used[k[1]] = used[k[2]] = 1;
			}
		}
	std::vector<int> remap(vm.v.size(), -1);
	SurfaceMesh sm;
	for (size_t i = 0; i < vm.v.size(); ++i)
		if (used[i]) {
			remap[i] = (int)sm.v.size();
			sm.v.push_back(vm.v[i]);
		}
	for (auto &t : tris)
		sm.tris.push_back({ remap[t[0]], remap[t[1]], remap[t[2]] });
	finish_surface(sm);
	return sm;
}

// ---- pressure fields (Drake make_*_field.h): E·extent, extent in [0,1], 0 on the boundary ------
static double snap(double extent) { return std::fabs(extent) < 1e-14 ? 0.0 : extent; }

std::vector<double> sphere_pressure(const VolumeMesh &vm, double r, double E)
{
	std::vector<double> e;
	for (auto &p : vm.v)
		e.push_back(E * snap(1.0 - norm(p) / r));
	return e;
}
std::vector<double> ellipsoid_pressure(const VolumeMesh &vm, double a, double b, double c, double E)
{
	std::vector<double> e;
	for (auto &p : vm.v)
		e.push_back(E * snap(1.0 - norm(V3{ p.x / a, p.y / b, p.z / c })));
	return e;
}
std::vector<double> box_pressure(const VolumeMesh &vm, double sx, double sy, double sz, double E)
{
	double h[3] = { sx / 2, sy / 2, sz / 2 };
	double m    = std::min(h[0], std::min(h[1], h[2]));
	std::vector<double> e;
	for (auto &p : vm.v) {
		double d = std::min(h[0] - std::fabs(p.x), std::min(h[1] - std::fabs(p.y), h[2] - std::fabs(p.z)));
		e.push_back(E * snap(d / m));
	}
	return e;
}
std::vector<double> cylinder_pressure(const VolumeMesh &vm, double r, double length, double E)
{
	double h = length / 2, m = std::min(r, h);
	std::vector<double> e;
	for (auto &p : vm.v) {
		double rho = std::sqrt(p.x * p.x + p.y * p.y);
		double d   = std::min(r - rho, h - std::fabs(p.z));
		e.push_back(E * snap(d / m));
	}
	return e;
}
std::vector<double> convex_pressure(const VolumeMesh &vm, double E)
{
	std::vector<double> e(vm.v.size(), 0.0);
	e.back() = E;
	return e;
}

// Drake VolumeMesh::CalcGradBarycentric / MeshFieldLinear::CalcValueAtMeshOrigin
VolumeField make_field(const VolumeMesh &vm, std::vector<double> e)
{
	VolumeField f;
	f.e = std::move(e);
	for (auto &t : vm.tets) {
		V3 g{ 0, 0, 0 };
		for (int i = 0; i < 4; ++i) {
			V3 pV = vm.v[t[i]], pA = vm.v[t[(i + 1) % 4]], pB = vm.v[t[(i + 2) % 4]], pC = vm.v[t[(i + 3) % 4]];
			V3 area_vec   = cross(pB - pA, pC - pA);
			double signed_volume = dot(area_vec, pV - pA);
			V3 gb         = area_vec / signed_volume;
			g             = i == 0 ? f.e[t[0]] * gb : g + f.e[t[i]] * gb;
		}
		f.grad.push_back(g);
		f.e0.push_back(f.e[t[0]] - dot(g, vm.v[t[0]]));
	}
	return f;
}

// ---- BVH: top-down median split on the longest axis, one element per leaf -----------------------
static void bounds_of(const std::vector<V3> &verts, const int *elems, int nper, const std::vector<int> &ids, int lo,
                      int hi, V3 &mn, V3 &mx)
{
	mn = { kInf, kInf, kInf };
	mx = { -kInf, -kInf, -kInf };
	for (int q = lo; q < hi; ++q)
		for (int k = 0; k < nper; ++k) {
			V3 p = verts[elems[ids[q] * nper + k]];
			mn   = { std::min(mn.x, p.x), std::min(mn.y, p.y), std::min(mn.z, p.z) };
			mx   = { std::max(mx.x, p.x), std::max(mx.y, p.y), std::max(mx.z, p.z) };
		}
}

static int build_rec(Bvh &bvh, const std::vector<V3> &verts, const int *elems, int nper, std::vector<int> &ids,
                     const std::vector<V3> &cent, int lo, int hi)
{
	int me = (int)bvh.nodes.size();
	bvh.nodes.emplace_back();
	V3 mn, mx;
	bounds_of(verts, elems, nper, ids, lo, hi, mn, mx);
	bvh.nodes[me].c = (mn + mx) * 0.5;
	// pad the half extents so the SAT stays conservative under rounding
	bvh.nodes[me].h = (mx - mn) * 0.5 + V3{ 1e-12, 1e-12, 1e-12 };
	if (hi - lo == 1) {
		bvh.nodes[me].elem = ids[lo];
		return me;
	}
	V3 ext   = mx - mn;
	int axis = ext.x >= ext.y && ext.x >= ext.z ? 0 : (ext.y >= ext.z ? 1 : 2);
	int mid  = (lo + hi) / 2;
	std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi,
	                 [&](int a, int b) { return cent[a][axis] < cent[b][axis] || (cent[a][axis] == cent[b][axis] && a < b); });
	int l                = build_rec(bvh, verts, elems, nper, ids, cent, lo, mid);
	int r                = build_rec(bvh, verts, elems, nper, ids, cent, mid, hi);
	bvh.nodes[me].left   = l;
	bvh.nodes[me].right  = r;
	return me;
}

Bvh build_bvh(const std::vector<V3> &verts, const int *elems, int nper, int nelem)
{
	Bvh bvh;
	if (nelem == 0)
		return bvh;
	std::vector<int> ids(nelem);
	std::vector<V3> cent(nelem);
	for (int i = 0; i < nelem; ++i) {
		ids[i] = i;
		V3 c{ 0, 0, 0 };
		for (int k = 0; k < nper; ++k)
			c = c + verts[elems[i * nper + k]];
		cent[i] = c / (double)nper;
	}
	bvh.nodes.reserve(2 * nelem);
	build_rec(bvh, verts, elems, nper, ids, cent, 0, nelem);
	return bvh;
}

} // namespace orc
