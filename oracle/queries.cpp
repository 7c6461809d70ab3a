// CPU ORACLE (test infrastructure, parity unpinned) — the three contact-surface queries.
//
// Restates Drake v1.8.0 geometry/proximity/{mesh_intersection.cc, mesh_plane_intersection.cc,
// field_intersection.cc, contact_surface_utility.cc, posed_half_space.h, plane.h,
// triangle_surface_mesh.h, polygon_surface_mesh.cc} and geometry/query_results/contact_surface.cc
// as reached from mujoco_contact_surfaces_plugin.cpp:284-303.  SURVEY.md App. A.3-A.7.
#include "oracle.hpp"

#include <algorithm>
#include <functional>

namespace orc {

namespace {

struct HalfSpace { // PosedHalfSpace / Plane: nhat·x − d
	V3 n;
	double d;
	HalfSpace(V3 normal, V3 point, bool already_normalized = false)
	{
		n = already_normalized ? normal : normalized(normal);
		d = dot(n, point);
	}
	double sd(V3 p) const { return dot(n, p) - d; }
};

typedef std::vector<V3> Poly;

// mesh_intersection.cc CalcIntersection
V3 calc_intersection(V3 A, V3 B, const HalfSpace &H)
{
	double a = H.sd(A), b = H.sd(B);
	double wa = b / (b - a);
	double wb = 1.0 - wa;
	return wa * A + wb * B;
}

// mesh_intersection.cc ClipPolygonByHalfSpace (Sutherland–Hodgman step)
void clip_polygon(const Poly &in, const HalfSpace &H, Poly &out)
{
	out.clear();
	int size = (int)in.size();
	for (int i = 0; i < size; ++i) {
		V3 current = in[i], previous = in[(i - 1 + size) % size];
		bool cin = H.sd(current) <= 0, pin = H.sd(previous) <= 0;
		if (cin) {
			if (!pin)
				out.push_back(calc_intersection(current, previous, H));
			out.push_back(current);
		} else if (pin) {
			out.push_back(calc_intersection(current, previous, H));
		}
	}
}

// mesh_intersection.cc RemoveDuplicateVertices / field_intersection.cc RemoveNearlyDuplicateVertices
void remove_duplicates(Poly &p)
{
	auto near = [](V3 a, V3 b) { return norm2(a - b) < 1e-14 * 1e-14; };
	p.erase(std::unique(p.begin(), p.end(), near), p.end());
	if (p.size() >= 3 && near(p.front(), p.back()))
		p.pop_back();
}

const int kTetFaces[4][3] = { { 1, 2, 3 }, { 0, 3, 2 }, { 0, 1, 3 }, { 0, 2, 1 } };

// mesh_plane_intersection.cc kTetEdges / kMarchingTetsTable.  Bit i of the code = vertex i is on
// the positive side.  Edge order makes the polygon's right-handed normal the plane normal.
const int kTetEdges[6][2]       = { { 0, 1 }, { 1, 2 }, { 2, 0 }, { 0, 3 }, { 1, 3 }, { 2, 3 } };
const int kMarchingTets[16][4]  = { { -1, -1, -1, -1 }, { 0, 3, 2, -1 }, { 0, 1, 4, -1 },  { 4, 3, 2, 1 },
	                                { 1, 2, 5, -1 },    { 0, 3, 5, 1 },  { 0, 2, 5, 4 },   { 3, 5, 4, -1 },
	                                { 3, 4, 5, -1 },    { 4, 5, 2, 0 },  { 1, 5, 3, 0 },   { 1, 5, 2, -1 },
	                                { 1, 2, 3, 4 },     { 0, 4, 1, -1 }, { 0, 2, 3, -1 },  { -1, -1, -1, -1 } };

// contact_surface_utility.cc CalcPolygonCentroid
V3 polygon_centroid(const std::vector<int> &poly, V3 n, const std::vector<V3> &v)
{
	int cnt = (int)poly.size();
	if (cnt == 3)
		return ((v[poly[0]] + v[poly[1]]) + v[poly[2]]) / 3.0;
	V3 acc{ 0, 0, 0 };
	double total = 0;
	V3 p0        = v[poly[0]];
	for (int i = 1; i < cnt - 1; ++i) {
		V3 p1 = v[poly[i]], p2 = v[poly[i + 1]];
		double a2 = dot(cross(p1 - p0, p2 - p0), n); // twice the signed fan-triangle area
		total += a2;
		acc = acc + a2 * ((p0 + p1) + p2);
	}
	return acc / (3.0 * total);
}

// TriMeshBuilder / PolyMeshBuilder (contact_surface_utility.h) writing into a Surface whose
// coordinates are still in the builder frame B; finish() maps to world.
struct Builder {
	Surface &s;
	std::vector<V3> face_grad_B; // per face, field gradient of the sampled field (frame B)
	explicit Builder(Surface &s_) : s(s_) {}
	int add_vertex(V3 p, double e)
	{
		s.v.push_back(p);
		s.e.push_back(e);
		return (int)s.v.size() - 1;
	}
	int add_polygon(const std::vector<int> &poly, V3 nhat, V3 grad)
	{
		int n = (int)poly.size();
		if (!s.tri) {
			s.face_first.push_back((int)s.face_idx.size());
			s.face_n.push_back(n);
			for (int i : poly)
				s.face_idx.push_back(i);
			face_grad_B.push_back(grad);
			return 1;
		}
		// AddPolygonToTriangleMeshData: centroid vertex + fan (prev, next, centroid)
		V3 c   = polygon_centroid(poly, nhat, s.v);
		int v0 = poly[0];
		double ec = s.e[v0] + dot(grad, c - s.v[v0]);
		int ci = add_vertex(c, ec);
		int cur = poly[n - 1];
		for (int i = 0; i < n; ++i) {
			int next = poly[i];
			s.face_first.push_back((int)s.face_idx.size());
			s.face_n.push_back(3);
			s.face_idx.push_back(cur);
			s.face_idx.push_back(next);
			s.face_idx.push_back(ci);
			face_grad_B.push_back(grad);
			cur = next;
		}
		return n;
	}
	// Mesh constructors (areas/normals/centroids in frame B), then TransformVertices(X_WB).
	void finish(const Xf &X_WB)
	{
		int F = s.num_faces();
		s.face_area.resize(F);
		s.face_normal.resize(F);
		s.face_centroid.resize(F);
		for (int f = 0; f < F; ++f) {
			const int *idx = &s.face_idx[s.face_first[f]];
			if (s.tri) { // TriangleSurfaceMesh::CalcAreasNormalsAndCentroid
				V3 a = s.v[idx[0]], b = s.v[idx[1]], c = s.v[idx[2]];
				V3 cr     = cross(b - a, c - a);
				double nn = norm(cr);
				s.face_area[f]   = 0.5 * nn;
				s.face_normal[f] = nn != 0.0 ? cr / nn : cr;
			} else { // PolygonSurfaceMesh::CalcAreaNormalAndCentroid, fan about vertex 0
				int n = s.face_n[f];
				V3 a  = s.v[idx[0]];
				double poly_area = 0;
				V3 nsum{ 0, 0, 0 }, csum{ 0, 0, 0 };
				for (int i = 1; i < n - 1; ++i) {
					V3 b = s.v[idx[i]], c = s.v[idx[i + 1]];
					V3 cr       = cross(b - a, c - a);
					double tri2 = norm(cr);
					poly_area += tri2;
					nsum = nsum + cr;
					csum = csum + tri2 * ((a + b) + c);
				}
				s.face_area[f]     = 0.5 * poly_area;
				s.face_normal[f]   = normalized(nsum);
				s.face_centroid[f] = poly_area != 0.0 ? csum / (3.0 * poly_area) : a;
			}
		}
		for (auto &p : s.v)
			p = apply(X_WB, p);
		for (int f = 0; f < F; ++f) {
			s.face_normal[f] = mul(X_WB.R, s.face_normal[f]);
			const int *idx   = &s.face_idx[s.face_first[f]];
			if (s.tri) // TriangleSurfaceMesh::element_centroid from the (world) vertices
				s.face_centroid[f] = ((s.v[idx[0]] + s.v[idx[1]]) + s.v[idx[2]]) / 3.0;
			else
				s.face_centroid[f] = apply(X_WB, s.face_centroid[f]);
		}
		if (!s.tri) { // MeshFieldLinear with given gradients; value at world origin from vertex 0
			s.poly_grad.resize(F);
			s.poly_e0.resize(F);
			for (int f = 0; f < F; ++f) {
				s.poly_grad[f] = mul(X_WB.R, face_grad_B[f]);
				int v0         = s.face_idx[s.face_first[f]];
				s.poly_e0[f]   = s.e[v0] - dot(s.poly_grad[f], s.v[v0]);
			}
		}
	}
};

// ContactSurface ctor: M is the geometry with the smaller id; swapping reverses face winding.
void order_ids(Surface &s, int gA, int gB)
{
	s.gM = gA;
	s.gN = gB;
	if (gB < gA) {
		std::swap(s.gM, s.gN);
		std::swap(s.has_gradM, s.has_gradN);
		s.gradM.swap(s.gradN);
		for (int f = 0; f < s.num_faces(); ++f) {
			int *idx = &s.face_idx[s.face_first[f]];
			if (s.tri)
				std::swap(idx[0], idx[1]);
			else
				std::reverse(idx + 1, idx + s.face_n[f]);
			s.face_normal[f] = -s.face_normal[f];
		}
		for (auto &em : s.emitted)
			std::swap(em.elemM, em.elemN);
	}
}

// ---- broadphase: dual BVH descent with a 15-axis SAT (Drake Bvh<Obb>::Collide) -------------------
bool boxes_overlap(const BvNode &a, const BvNode &b, const Xf &X_AB)
{
	// b's box expressed in a's frame: centre t, axes = columns of R
	const M3 &R = X_AB.R;
	V3 t        = apply(X_AB, b.c) - a.c;
	double AR[9];
	for (int i = 0; i < 9; ++i)
		AR[i] = std::fabs(R.m[i]) + 1e-14;
	double ah[3] = { a.h.x, a.h.y, a.h.z }, bh[3] = { b.h.x, b.h.y, b.h.z }, tt[3] = { t.x, t.y, t.z };
	for (int i = 0; i < 3; ++i) {
		double rb = bh[0] * AR[3 * i] + bh[1] * AR[3 * i + 1] + bh[2] * AR[3 * i + 2];
		if (std::fabs(tt[i]) > ah[i] + rb)
			return false;
	}
	for (int j = 0; j < 3; ++j) {
		double ra = ah[0] * AR[j] + ah[1] * AR[3 + j] + ah[2] * AR[6 + j];
		double tj = tt[0] * R.m[j] + tt[1] * R.m[3 + j] + tt[2] * R.m[6 + j];
		if (std::fabs(tj) > ra + bh[j])
			return false;
	}
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
			double ra = ah[i1] * AR[3 * i2 + j] + ah[i2] * AR[3 * i1 + j];
			double rb = bh[j1] * AR[3 * i + j2] + bh[j2] * AR[3 * i + j1];
			double tv = tt[i2] * R.m[3 * i1 + j] - tt[i1] * R.m[3 * i2 + j];
			if (std::fabs(tv) > ra + rb)
				return false;
		}
	return true;
}

void collide(const Bvh &A, const Bvh &B, const Xf &X_AB, const std::function<void(int, int)> &cb)
{
	if (A.nodes.empty() || B.nodes.empty())
		return;
	std::vector<std::pair<int, int>> stack = { { 0, 0 } };
	while (!stack.empty()) {
		auto [ia, ib] = stack.back();
		stack.pop_back();
		const BvNode &a = A.nodes[ia], &b = B.nodes[ib];
		if (!boxes_overlap(a, b, X_AB))
			continue;
		bool la = a.elem >= 0, lb = b.elem >= 0;
		if (la && lb) {
			cb(a.elem, b.elem);
		} else if (la) {
			stack.push_back({ ia, b.right });
			stack.push_back({ ia, b.left });
		} else if (lb) {
			stack.push_back({ a.right, ib });
			stack.push_back({ a.left, ib });
		} else {
			stack.push_back({ a.right, b.right });
			stack.push_back({ a.right, b.left });
			stack.push_back({ a.left, b.right });
			stack.push_back({ a.left, b.left });
		}
	}
}

void collide_plane(const Bvh &A, V3 n, double d, const std::function<void(int)> &cb)
{
	if (A.nodes.empty())
		return;
	std::vector<int> stack = { 0 };
	while (!stack.empty()) {
		int ia = stack.back();
		stack.pop_back();
		const BvNode &a = A.nodes[ia];
		double dist     = dot(n, a.c) - d;
		double rad      = a.h.x * std::fabs(n.x) + a.h.y * std::fabs(n.y) + a.h.z * std::fabs(n.z);
		if (std::fabs(dist) > rad)
			continue; // box entirely on one side: no tet of it is cut
		if (a.elem >= 0)
			cb(a.elem);
		else {
			stack.push_back(a.right);
			stack.push_back(a.left);
		}
	}
}

} // namespace

// mesh_intersection.cc ComputeContactSurfaceFromSoftVolumeRigidSurface (plugin.cpp:301-303)
std::shared_ptr<Surface> soft_rigid(const Geom &S, int gS, const Xf &X_WS, const Geom &R, int gR, const Xf &X_WR,
                                    bool tri, bool use_bvh)
{
	auto s  = std::make_shared<Surface>();
	s->tri  = tri;
	Xf X_SR = invert_and_compose(X_WS, X_WR);
	Builder builder(*s);
	Poly poly, tmp;
	std::vector<int> idx;
	std::vector<V3> gradS;
	auto cb = [&](int tet, int f) {
		s->n_candidates++;
		// IsFaceNormalAlongPressureGradient
		V3 ghat   = normalized(S.pf.grad[tet]);
		V3 nhat_S = mul(X_SR.R, R.sm.normal[f]);
		if (!(dot(ghat, nhat_S) > kCosAlpha))
			return;
		// ClipTriangleByTetrahedron
		poly.clear();
		for (int i = 0; i < 3; ++i)
			poly.push_back(apply(X_SR, R.sm.v[R.sm.tris[f][i]]));
		V3 pv[4];
		for (int i = 0; i < 4; ++i)
			pv[i] = S.vm.v[S.vm.tets[tet][i]];
		for (auto &fv : kTetFaces) {
			V3 A = pv[fv[0]], B = pv[fv[1]], C = pv[fv[2]];
			HalfSpace H(cross(B - A, C - A), A);
			clip_polygon(poly, H, tmp);
			poly.swap(tmp);
		}
		remove_duplicates(poly);
		if (poly.size() < 3)
			return;
		idx.clear();
		for (auto &p : poly)
			idx.push_back(builder.add_vertex(p, dot(S.pf.grad[tet], p) + S.pf.e0[tet]));
		int first = s->num_faces();
		int nf    = builder.add_polygon(idx, nhat_S, S.pf.grad[tet]);
		for (int i = 0; i < nf; ++i)
			gradS.push_back(S.pf.grad[tet]);
		s->emitted.push_back({ tet, f, (int)poly.size(), first, nf });
	};
	if (use_bvh)
		collide(S.bvh, R.bvh, X_SR, cb);
	else
		for (int t = 0; t < (int)S.vm.tets.size(); ++t)
			for (int f = 0; f < (int)R.sm.tris.size(); ++f)
				cb(t, f);
	if (s->num_faces() == 0)
		return nullptr;
	builder.finish(X_WS);
	s->has_gradM = true;
	for (auto &g : gradS)
		s->gradM.push_back(mul(X_WS.R, g));
	order_ids(*s, gS, gR);
	return s;
}

// mesh_plane_intersection.cc ComputeContactSurfaceFromSoftVolumeRigidHalfSpace (plugin.cpp:298-299)
std::shared_ptr<Surface> soft_plane(const Geom &S, int gS, const Xf &X_WS, int gR, const Xf &X_WR, bool tri,
                                    bool use_bvh)
{
	auto s  = std::make_shared<Surface>();
	s->tri  = tri;
	Xf X_SR = invert_and_compose(X_WS, X_WR);
	V3 Rz_S = X_SR.R.col(2);
	HalfSpace plane(Rz_S, X_SR.p, true);
	Builder builder(*s);
	std::vector<V3> gradS;
	std::vector<int> idx;
	V3 nhat_W = mul(X_WS.R, plane.n);
	// Builder frame is the WORLD frame here (SliceTetWithPlane adds X_WM * p_MC).
	auto cb = [&](int tet) {
		s->n_candidates++;
		const auto &T = S.vm.tets[tet];
		double dist[4];
		int code = 0;
		for (int i = 0; i < 4; ++i) {
			dist[i] = plane.sd(S.vm.v[T[i]]);
			if (dist[i] > 0)
				code |= 1 << i;
		}
		const int *edges = kMarchingTets[code];
		if (edges[0] == -1)
			return;
		idx.clear();
		for (int e = 0; e < 4 && edges[e] != -1; ++e) {
			int l0 = kTetEdges[edges[e]][0], l1 = kTetEdges[edges[e]][1];
			// canonical edge direction (lower global vertex id first): equals Drake's cached cut vertex
			// whichever tet reaches the shared edge first, up to that tet's edge orientation
			if (T[l0] > T[l1])
				std::swap(l0, l1);
			V3 p0 = S.vm.v[T[l0]], p1 = S.vm.v[T[l1]];
			double t  = dist[l0] / (dist[l0] - dist[l1]);
			V3 pc     = p0 + t * (p1 - p0);
			double e0 = S.pf.e[T[l0]], e1 = S.pf.e[T[l1]];
			idx.push_back(builder.add_vertex(apply(X_WS, pc), e0 + t * (e1 - e0)));
		}
		V3 grad_W = mul(X_WS.R, S.pf.grad[tet]);
		int first = s->num_faces();
		int nv    = (int)idx.size();
		int nf    = builder.add_polygon(idx, nhat_W, grad_W);
		for (int i = 0; i < nf; ++i)
			gradS.push_back(grad_W);
		s->emitted.push_back({ tet, 0, nv, first, nf });
	};
	if (use_bvh)
		collide_plane(S.bvh, plane.n, plane.d, cb);
	else
		for (int t = 0; t < (int)S.vm.tets.size(); ++t)
			cb(t);
	if (s->num_faces() == 0)
		return nullptr;
	Xf I{ { { 1, 0, 0, 0, 1, 0, 0, 0, 1 } }, { 0, 0, 0 } };
	builder.finish(I);
	s->has_gradM = true;
	s->gradM     = gradS;
	order_ids(*s, gS, gR);
	return s;
}

// field_intersection.cc ComputeContactSurfaceFromCompliantVolumes (plugin.cpp:284-286)
std::shared_ptr<Surface> soft_soft(const Geom &A, int gA, const Xf &X_WA, const Geom &B, int gB, const Xf &X_WB,
                                   bool tri, bool use_bvh)
{
	auto s  = std::make_shared<Surface>();
	s->tri  = tri;
	Xf X_MN = invert_and_compose(X_WA, X_WB);
	Builder builder(*s);
	std::vector<V3> g0s, g1s;
	Poly poly, tmp;
	std::vector<int> idx;
	// p_NMo = X_MN^-1 translation
	V3 p_NMo = -mulT(X_MN.R, X_MN.p);
	auto cb  = [&](int t0, int t1) {
		s->n_candidates++;
		// CalcEquilibriumPlane
		V3 grad0     = A.pf.grad[t0];
		double f0_Mo = A.pf.e0[t0]; // EvaluateCartesian(tet0, 0)
		V3 grad1_N   = B.pf.grad[t1];
		V3 grad1_M   = mul(X_MN.R, grad1_N);
		double f1_Mo = dot(grad1_N, p_NMo) + B.pf.e0[t1];
		V3 n_M       = grad0 - grad1_M;
		double mag   = norm(n_M);
		if (mag <= 0.0)
			return;
		V3 nhat_M = n_M / mag;
		V3 p_MQ   = -((f0_Mo - f1_Mo) / mag) * nhat_M;
		HalfSpace plane(nhat_M, p_MQ, true);
		// IsPlaneNormalAlongPressureGradient for both fields
		if (!(dot(nhat_M, normalized(grad0)) > kCosAlpha))
			return;
		V3 rev_N = mulT(X_MN.R, -nhat_M);
		if (!(dot(rev_N, normalized(grad1_N)) > kCosAlpha))
			return;
		// IntersectTetrahedra: slice tet0 with the plane ...
		const auto &T0 = A.vm.tets[t0];
		double dist[4];
		int code = 0;
		for (int i = 0; i < 4; ++i) {
			dist[i] = plane.sd(A.vm.v[T0[i]]);
			if (dist[i] > 0)
				code |= 1 << i;
		}
		poly.clear();
		const int *edges = kMarchingTets[code];
		for (int e = 0; e < 4 && edges[e] != -1; ++e) {
			int l0 = kTetEdges[edges[e]][0], l1 = kTetEdges[edges[e]][1];
			V3 p0 = A.vm.v[T0[l0]], p1 = A.vm.v[T0[l1]];
			double t = dist[l0] / (dist[l0] - dist[l1]);
			poly.push_back(p0 + t * (p1 - p0));
		}
		remove_duplicates(poly);
		if (poly.size() < 3)
			return;
		// ... then clip by the four half spaces of tet1 expressed in M
		V3 pv[4];
		for (int i = 0; i < 4; ++i)
			pv[i] = apply(X_MN, B.vm.v[B.vm.tets[t1][i]]);
		for (auto &fv : kTetFaces) {
			V3 PA = pv[fv[0]], PB = pv[fv[1]], PC = pv[fv[2]];
			HalfSpace H(cross(PB - PA, PC - PA), PA);
			clip_polygon(poly, H, tmp);
			remove_duplicates(tmp);
			if (tmp.size() < 3)
				return;
			poly.swap(tmp);
		}
		idx.clear();
		for (auto &p : poly)
			idx.push_back(builder.add_vertex(p, dot(grad0, p) + f0_Mo));
		int first = s->num_faces();
		int nf    = builder.add_polygon(idx, nhat_M, grad0);
		for (int i = 0; i < nf; ++i) {
			g0s.push_back(grad0);
			g1s.push_back(grad1_M);
		}
		s->emitted.push_back({ t0, t1, (int)poly.size(), first, nf });
	};
	if (use_bvh)
		collide(A.bvh, B.bvh, X_MN, cb);
	else
		for (int a = 0; a < (int)A.vm.tets.size(); ++a)
			for (int b = 0; b < (int)B.vm.tets.size(); ++b)
				cb(a, b);
	if (s->num_faces() == 0)
		return nullptr;
	builder.finish(X_WA);
	s->has_gradM = s->has_gradN = true;
	for (auto &g : g0s)
		s->gradM.push_back(mul(X_WA.R, g));
	for (auto &g : g1s)
		s->gradN.push_back(mul(X_WA.R, g));
	order_ids(*s, gA, gB);
	return s;
}

} // namespace orc
