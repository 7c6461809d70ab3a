// CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY: tactile ray caster PINNED to the reference's compiled code;
// Drake share UNPINNED.**
//
// This directory is a dependency-free C++17/fp64 restatement of the hot path of
// ubi-agni/mujoco_contact_surfaces: the Drake v1.8.0 `geometry/proximity` arithmetic the
// plugin calls (un-vendored third-party dependency pinned by the reference's README.md:7-10,
// release drake-20220919), plus the arithmetic the reference owns itself
// (mujoco_contact_surfaces/src/mujoco_contact_surfaces_plugin.cpp:107-187, 255-523, 571-813 and
// mujoco_contact_surface_sensors/src/{flat_tactile_sensor.cpp:127-214,262-402, bvh.cpp:49-476}).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (mujoco_contact_surfaces_b200/) never links, imports or calls it.
//
// What pins it:
// * PINNED — the float32 BVH / TLAS / Moeller-Trumbore / slab-test ray caster under the flat and curved sensors
//   (sensor.cpp: Blas, Tlas, intersect_triangle, intersect_aabb): the reference's own bvh.cpp + bvh.h + float3.h
//   compile unmodified against container-only shim headers (oracle/ref_shim) into oracle/_ref/, and
//   tests/test_ref_pinning.py holds this restatement to that compiled code bit for bit — live, and through the vectors
//   scripts/make_ref_golden.py made with it (tests/golden/ref_bvh_vectors.npz).
// * UNPINNED — everything Drake computes (meshes, fields, the three contact queries, ContactSurface) and the plugin's
//   force loop: the reference has no tests and no golden vectors, and Drake / MuJoCo / ROS cannot be built or imported in
//   this environment, so that share is checked against analytic known-answer tests
//   (tests/test_oracle_kat.py) and not against reference outputs.
//
// Build: `make -C oracle` (g++ -O2 -ffp-contract=off; x86-64 SSE2 IEEE double, no FMA contraction).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Fixed-order 3-vector algebra.  dot() uses Eigen's unrolled reduction tree
// for fixed size 3: a0*b0 + (a1*b1 + a2*b2).  No FMA contraction anywhere (-ffp-contract=off).
// ---------------------------------------------------------------------------------------------
struct V3 {
	double x, y, z;
	double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	double &at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator-(V3 a) { return { -a.x, -a.y, -a.z }; }
inline V3 operator*(V3 a, double s) { return { a.x * s, a.y * s, a.z * s }; }
inline V3 operator*(double s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline V3 operator/(V3 a, double s) { return { a.x / s, a.y / s, a.z / s }; }
inline double dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline double norm2(V3 a) { return dot(a, a); }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
// Eigen::normalized(): z = squaredNorm; z > 0 ? v / sqrt(z) : v
inline V3 normalized(V3 a)
{
	double z = dot(a, a);
	return z > 0 ? a / std::sqrt(z) : a;
}

struct M3 { // row-major
	double m[9];
	V3 row(int i) const { return { m[3 * i], m[3 * i + 1], m[3 * i + 2] }; }
	V3 col(int j) const { return { m[j], m[3 + j], m[6 + j] }; }
};
inline V3 mul(const M3 &R, V3 v) { return { dot(R.row(0), v), dot(R.row(1), v), dot(R.row(2), v) }; }
inline V3 mulT(const M3 &R, V3 v) { return { dot(R.col(0), v), dot(R.col(1), v), dot(R.col(2), v) }; }
inline M3 mulTM(const M3 &A, const M3 &B) // A^T * B
{
	M3 C;
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			C.m[3 * i + j] = dot(A.col(i), B.col(j));
	return C;
}

struct Xf { // rigid transform X_AB: p_A = R * p_B + p
	M3 R;
	V3 p;
};
inline V3 apply(const Xf &X, V3 v) { return mul(X.R, v) + X.p; }
// drake::math::RigidTransform::InvertAndCompose: X_AC = X_BA^-1 * X_BC
inline Xf invert_and_compose(const Xf &X_BA, const Xf &X_BC)
{
	Xf X;
	X.R = mulTM(X_BA.R, X_BC.R);
	X.p = mulT(X_BA.R, X_BC.p - X_BA.p);
	return X;
}

// cos(5*pi/8): Drake mesh_intersection.cc / field_intersection.cc kAlpha = 5π/8 (glibc value).
constexpr double kCosAlpha = -0x1.87de2a6aea962p-2;
constexpr double kInf      = std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------------------------------------
// Meshes and fields
// ---------------------------------------------------------------------------------------------
struct VolumeMesh {
	std::vector<V3> v;
	std::vector<std::array<int, 4>> tets;
};
struct SurfaceMesh {
	std::vector<V3> v;
	std::vector<std::array<int, 3>> tris;
	std::vector<V3> normal; // unit face normals
	std::vector<double> area;
};
struct VolumeField { // drake VolumeMeshFieldLinear<double,double>
	std::vector<double> e; // per vertex
	std::vector<V3> grad; // per tet
	std::vector<double> e0; // per tet value at mesh-frame origin
};

int sphere_refinement_level(double r, double hint);
VolumeMesh make_unit_sphere_volume(int level);
VolumeMesh make_sphere_volume(double r, double hint);
VolumeMesh make_ellipsoid_volume(double a, double b, double c, double hint);
VolumeMesh make_box_volume(double sx, double sy, double sz, double hint); // full sizes
VolumeMesh make_box_volume_ma(double sx, double sy, double sz);
VolumeMesh make_cylinder_volume_ma(double r, double length, double hint);
VolumeMesh make_convex_volume(const SurfaceMesh &sm); // plugin.cpp:161-187, 767-787
SurfaceMesh volume_to_surface(const VolumeMesh &vm);
void finish_surface(SurfaceMesh &sm); // normals + areas

std::vector<double> sphere_pressure(const VolumeMesh &vm, double r, double E);
std::vector<double> ellipsoid_pressure(const VolumeMesh &vm, double a, double b, double c, double E);
std::vector<double> box_pressure(const VolumeMesh &vm, double sx, double sy, double sz, double E);
std::vector<double> cylinder_pressure(const VolumeMesh &vm, double r, double length, double E);
std::vector<double> convex_pressure(const VolumeMesh &vm, double E);
VolumeField make_field(const VolumeMesh &vm, std::vector<double> e);

// ---------------------------------------------------------------------------------------------
// BVH used by the oracle's broadphase (Drake Bvh<Obb,·>::Collide restated with local-frame
// AABBs and a 15-axis SAT).  Candidate sets are BV specific; parity is
// defined on the EMITTED set.
// ---------------------------------------------------------------------------------------------
struct BvNode {
	V3 c, h; // centre, half extents (local frame)
	int left = -1, right = -1; // children, or
	int elem = -1; // leaf element
};
struct Bvh {
	std::vector<BvNode> nodes; // root = 0
};
Bvh build_bvh(const std::vector<V3> &verts, const int *elems, int nper, int nelem);

// ---------------------------------------------------------------------------------------------
// Contact surface (drake::geometry::ContactSurface<double> restated)
// ---------------------------------------------------------------------------------------------
struct Emitted { // provenance of one polygon
	int elemM, elemN, nverts, first_face, n_faces;
};
struct Surface {
	bool tri = false; // kTriangle vs kPolygon
	int gM = -1, gN = -1; // config indices after the id-ordering swap
	std::vector<V3> v; // world frame
	std::vector<double> e; // per-vertex pressure
	std::vector<int> face_first, face_n; // index into face_idx
	std::vector<int> face_idx;
	std::vector<V3> face_normal, face_centroid;
	std::vector<double> face_area;
	bool has_gradM = false, has_gradN = false;
	std::vector<V3> gradM, gradN; // world frame, per face
	std::vector<V3> poly_grad; // kPolygon field: per-face gradient (world) ...
	std::vector<double> poly_e0; // ... and value at world origin
	std::vector<Emitted> emitted;
	long n_candidates = 0;
	int num_faces() const { return (int)face_first.size(); }
};

struct PointCollision { // common_types.h:48-56
	V3 p, n;
	double fn0, stiffness, damping;
	int face;
};

enum GeomKind { RIGID_MESH = 0, SOFT = 1, RIGID_PLANE = 2 };

struct Geom { // ContactProperties, mujoco_contact_surfaces_plugin.h:141-213
	int kind = RIGID_MESH;
	int mj_type = 0;
	double E = kInf, dissipation = 1.0, mu_s = 0, mu_d = 0, hint = 0;
	VolumeMesh vm;
	VolumeField pf;
	SurfaceMesh sm;
	Bvh bvh;
};

struct FlatSensor {
	int geom;
	double resolution;
	int S;
	int window; // 0 none 1 gauss 2 tukey 3 square
	float sigma;
	double size[3]; // geom_size of the sensor geom
	int cx, cy;
};

// CurvedSensor (mujoco_contact_surface_sensors/src/curved_sensor.cpp): taxels with optional normals, surface sample
// points with normals (the reference draws them with vcglib's Poisson-disk sampler, seed 42 — absent here, so the
// samples are an INPUT), the per-taxel sample lists and weights of load() :325-368, rays cast in internal_update
// :388-481.
struct CurvedSensor {
	int geom;
	double include_margin;
	std::vector<V3> taxel_pos, taxel_nrm;        // geom frame; a zero normal disables the 45 degree test
	std::vector<V3> surf_pos, surf_nrm;          // "close" sample points (kept in first-use order, :352-356)
	std::vector<int> surf_src;                   // index of each close point in the caller's sample array
	std::vector<std::vector<int>> surface_idx;   // per taxel: indices into surf_pos
	std::vector<std::vector<double>> surface_weight;
};

// TaxelSensor (mujoco_contact_surface_sensors/src/taxel_sensor.cpp), sample_method "default" (:173-213): a
// barycentric lattice on every contact-surface triangle, taxel values from the samples within include_margin.
struct TaxelSensor {
	int geom;
	double include_margin, sample_resolution;
	int method;      // 0 closest, 1 weighted, 2 mean, 3 squared (:79-91)
	bool visualize;  // changes the VALUE of the closest method (:312-328), quirk Q12
	int sample_method = 0; // 0 DEFAULT (barycentric lattice, :174-209), 1 AREA_IMPORTANCE (:211-254)
	std::vector<V3> taxels; // geom frame
};

struct PairOut {
	bool has_surface = false;
	int gM = -1, gN = -1;
	std::shared_ptr<Surface> s;
	std::vector<PointCollision> pcs;
	V3 F{ 0, 0, 0 }, tau{ 0, 0, 0 }, centroid{ 0, 0, 0 };
	double area = 0;
	std::vector<V3> face_force; // per PointCollision
};

struct StepState { // everything one env step produces (geomCollisions + applied wrenches)
	std::vector<PairOut> out;
	std::vector<std::array<double, 6>> geom_wrench; // F, tau about world origin
	std::vector<double> xpos, xmat, vel;
	long n_candidates = 0;
};

struct Scene {
	bool tri = false;
	bool apply_forces = true;
	std::vector<Geom> geoms;
	std::vector<std::array<int, 2>> pairs;
	std::vector<FlatSensor> sensors;
	std::vector<CurvedSensor> curved;
	std::vector<TaxelSensor> taxel;
	StepState last;
};

// the three queries; use_bvh==false enumerates the Cartesian product
std::shared_ptr<Surface> soft_rigid(const Geom &S, int gS, const Xf &X_WS, const Geom &R, int gR, const Xf &X_WR,
                                    bool tri, bool use_bvh);
std::shared_ptr<Surface> soft_plane(const Geom &S, int gS, const Xf &X_WS, int gR, const Xf &X_WR, bool tri,
                                    bool use_bvh);
std::shared_ptr<Surface> soft_soft(const Geom &A, int gA, const Xf &X_WA, const Geom &B, int gB, const Xf &X_WB,
                                   bool tri, bool use_bvh);

void evaluate_contact_surface(const Scene &sc, PairOut &po); // plugin.cpp:320-409
void passive_forces(const Scene &sc, PairOut &po, const double *xpos, const double *vel); // plugin.cpp:411-483
void step(const Scene &sc, StepState &st, const double *xpos, const double *xmat, const double *vel, bool use_bvh);
// caster: 0 = linear scan over all triangles, 1 = the oracle's own restatement of the reference's BVH/TLAS,
// 2 = the reference's compiled ray caster (oracle/_ref) installed with set_external_caster.  trace (optional): every
// ray (O, D) in (x, y, i, j) order with its nearest hit (t, u, v, (blas << 20) + triangle).
struct FlatTrace {
	std::vector<float> rays, tuv;
	std::vector<uint32_t> id;
};
struct ExternalCaster { // signatures of oracle/ref_shim/ref_capi.cpp
	void *(*create)(int n_surf, const int *n_tri, const double *verts, const double *press);
	void (*cast)(void *h, int n, const float *O, const float *D, float *tuv, uint32_t *id);
	void (*destroy)(void *h);
};
void set_external_caster(const ExternalCaster &c);
void flat_sensor_image(const Scene &sc, const StepState &st, int sensor, float *out, int caster, bool parallel,
                       FlatTrace *trace = nullptr);
void cast_rays(int n_surf, const int *n_tri, const double *verts, int n_rays, const float *O, const float *D, float *tuv,
               uint32_t *id);
void intersect_triangle_one(const float *O, const float *D, const float *v0, const float *v1, const float *v2, float t_in,
                            float *tuv_out, int *hit_out);
float intersect_aabb_one(const float *O, const float *D, float t_in, const float *bmin, const float *bmax);
void curved_sensor_load(CurvedSensor &cs, const double *sample_pos, const double *sample_nrm, int n_samples);
void curved_sensor_values(const Scene &sc, const StepState &st, int sensor, float *out, int caster); // caster as above
// values: the message of the previous update on entry (taxels without a sample in range keep it), updated in place
void taxel_sensor_values(const Scene &sc, const StepState &st, int sensor, float *values);

double combined_dissipation(const Geom &a, const Geom &b);
double combined_friction_dynamic(const Geom &a, const Geom &b);

} // namespace orc
