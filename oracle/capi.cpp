// CPU ORACLE (test infrastructure; ray caster pinned to oracle/_ref, the rest unpinned: oracle.hpp) — C entry points for ctypes (oracle/oracle.py).
// Geometry configuration mirrors parseMujocoCustomFields (plugin.cpp:613-812): one call per
// `cs::<geom>` numeric in configuration order; the returned index plays the role of drake_id.
#include "oracle.hpp"

#include <chrono>
#include <cstring>
#include <string>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {
thread_local std::string g_err;
enum { mjGEOM_PLANE = 0, mjGEOM_HFIELD, mjGEOM_SPHERE, mjGEOM_CAPSULE, mjGEOM_ELLIPSOID, mjGEOM_CYLINDER, mjGEOM_BOX, mjGEOM_MESH };

void finish_geom(Geom &g)
{
	if (g.kind == SOFT) {
		g.bvh = build_bvh(g.vm.v, &g.vm.tets[0][0], 4, (int)g.vm.tets.size());
	} else if (g.kind == RIGID_MESH) {
		if (g.sm.normal.empty())
			finish_surface(g.sm);
		g.bvh = build_bvh(g.sm.v, &g.sm.tris[0][0], 3, (int)g.sm.tris.size());
	}
}
} // namespace

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void *orc_scene_create(int triangle_representation, int apply_forces)
{
	Scene *s        = new Scene();
	s->tri          = triangle_representation != 0;
	s->apply_forces = apply_forces != 0;
	return s;
}
void orc_scene_destroy(void *h) { delete (Scene *)h; }

// props = [hydroelasticModulus, dissipation, resolutionHint, staticFriction, dynamicFriction]
int orc_add_geom(void *h, int mj_type, const double *size, const float *mesh_vert, int nvert, const int *mesh_face,
                 int nface, const double *props)
{
	Scene &sc = *(Scene *)h;
	try {
		Geom g;
		g.mj_type = mj_type;
		double E = props[0], hint = props[2];
		bool soft = E > 0;
		g.mu_s    = props[3];
		g.mu_d    = props[4];
		g.hint    = hint;
		if (soft) {
			g.kind        = SOFT;
			g.E           = E;
			g.dissipation = props[1];
		}
		switch (mj_type) {
			case mjGEOM_PLANE:
				if (soft) {
					g_err = "soft plane collision not implemented (plugin.cpp:635-636)";
					return -1;
				}
				g.kind = RIGID_PLANE;
				break;
			case mjGEOM_SPHERE:
				g.vm = make_sphere_volume(size[0], hint);
				if (soft)
					g.pf = make_field(g.vm, sphere_pressure(g.vm, size[0], E));
				else
					g.sm = volume_to_surface(g.vm);
				break;
			case mjGEOM_ELLIPSOID:
				g.vm = make_ellipsoid_volume(size[0], size[1], size[2], hint);
				if (soft)
					g.pf = make_field(g.vm, ellipsoid_pressure(g.vm, size[0], size[1], size[2], E));
				else
					g.sm = volume_to_surface(g.vm);
				break;
			case mjGEOM_CYLINDER:
				g.vm = make_cylinder_volume_ma(size[0], 2 * size[1], hint);
				if (soft)
					g.pf = make_field(g.vm, cylinder_pressure(g.vm, size[0], 2 * size[1], E));
				else
					g.sm = volume_to_surface(g.vm);
				break;
			case mjGEOM_BOX: {
				double sx = 2 * size[0], sy = 2 * size[1], sz = 2 * size[2];
				if (soft) {
					g.vm = hint > 0 ? make_box_volume(sx, sy, sz, hint) : make_box_volume_ma(sx, sy, sz);
					g.pf = make_field(g.vm, box_pressure(g.vm, sx, sy, sz, E));
				} else {
					g.vm = make_box_volume(sx, sy, sz, hint);
					g.sm = volume_to_surface(g.vm);
				}
				break;
			}
			case mjGEOM_MESH: {
				if (nvert <= 0 || nface <= 0) {
					g_err = "Could not load mesh (plugin.cpp:804-805)";
					return -1;
				}
				for (int i = 0; i < nvert; ++i)
					g.sm.v.push_back({ (double)mesh_vert[3 * i], (double)mesh_vert[3 * i + 1], (double)mesh_vert[3 * i + 2] });
				for (int i = 0; i < nface; ++i)
					g.sm.tris.push_back({ mesh_face[3 * i], mesh_face[3 * i + 1], mesh_face[3 * i + 2] });
				finish_surface(g.sm);
				if (soft) {
					g.vm = make_convex_volume(g.sm);
					g.pf = make_field(g.vm, convex_pressure(g.vm, E));
				}
				break;
			}
			default:
				g_err = "geom type not implemented (hfield/capsule: plugin.cpp:642-647)";
				return -1;
		}
		if (!soft)
			g.vm = VolumeMesh();
		finish_geom(g);
		sc.geoms.push_back(std::move(g));
		return (int)sc.geoms.size() - 1;
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
}

// raw meshes for known-answer tests
int orc_add_raw_soft(void *h, const double *verts, int nv, const int *tets, int nt, const double *pressure,
                     const double *props)
{
	Scene &sc = *(Scene *)h;
	Geom g;
	g.kind = SOFT;
	g.mj_type = mjGEOM_MESH;
	g.E = props[0];
	g.dissipation = props[1];
	g.hint = props[2];
	g.mu_s = props[3];
	g.mu_d = props[4];
	for (int i = 0; i < nv; ++i)
		g.vm.v.push_back({ verts[3 * i], verts[3 * i + 1], verts[3 * i + 2] });
	for (int i = 0; i < nt; ++i)
		g.vm.tets.push_back({ tets[4 * i], tets[4 * i + 1], tets[4 * i + 2], tets[4 * i + 3] });
	g.pf = make_field(g.vm, std::vector<double>(pressure, pressure + nv));
	finish_geom(g);
	sc.geoms.push_back(std::move(g));
	return (int)sc.geoms.size() - 1;
}
int orc_add_raw_rigid(void *h, const double *verts, int nv, const int *tris, int nt, const double *props)
{
	Scene &sc = *(Scene *)h;
	Geom g;
	g.kind = RIGID_MESH;
	g.mj_type = mjGEOM_MESH;
	g.hint = props[2];
	g.mu_s = props[3];
	g.mu_d = props[4];
	for (int i = 0; i < nv; ++i)
		g.sm.v.push_back({ verts[3 * i], verts[3 * i + 1], verts[3 * i + 2] });
	for (int i = 0; i < nt; ++i)
		g.sm.tris.push_back({ tris[3 * i], tris[3 * i + 1], tris[3 * i + 2] });
	finish_geom(g);
	sc.geoms.push_back(std::move(g));
	return (int)sc.geoms.size() - 1;
}

// info = [kind, n_vertices, n_elements]
int orc_geom_info(void *h, int gi, int *info)
{
	Scene &sc = *(Scene *)h;
	if (gi < 0 || gi >= (int)sc.geoms.size())
		return -1;
	const Geom &g = sc.geoms[gi];
	info[0]       = g.kind;
	info[1]       = g.kind == SOFT ? (int)g.vm.v.size() : (int)g.sm.v.size();
	info[2]       = g.kind == SOFT ? (int)g.vm.tets.size() : (int)g.sm.tris.size();
	return 0;
}
// verts[nv*3], elems[ne*(4|3)], soft: pressure[nv], grad[ne*3], e0[ne]; rigid: normal->grad[ne*3]
int orc_geom_mesh(void *h, int gi, double *verts, int *elems, double *pressure, double *grad, double *e0)
{
	Scene &sc     = *(Scene *)h;
	const Geom &g = sc.geoms[gi];
	if (g.kind == SOFT) {
		for (size_t i = 0; i < g.vm.v.size(); ++i) {
			verts[3 * i] = g.vm.v[i].x, verts[3 * i + 1] = g.vm.v[i].y, verts[3 * i + 2] = g.vm.v[i].z;
			if (pressure)
				pressure[i] = g.pf.e[i];
		}
		for (size_t i = 0; i < g.vm.tets.size(); ++i) {
			for (int k = 0; k < 4; ++k)
				elems[4 * i + k] = g.vm.tets[i][k];
			if (grad)
				grad[3 * i] = g.pf.grad[i].x, grad[3 * i + 1] = g.pf.grad[i].y, grad[3 * i + 2] = g.pf.grad[i].z;
			if (e0)
				e0[i] = g.pf.e0[i];
		}
	} else if (g.kind == RIGID_MESH) {
		for (size_t i = 0; i < g.sm.v.size(); ++i)
			verts[3 * i] = g.sm.v[i].x, verts[3 * i + 1] = g.sm.v[i].y, verts[3 * i + 2] = g.sm.v[i].z;
		for (size_t i = 0; i < g.sm.tris.size(); ++i) {
			for (int k = 0; k < 3; ++k)
				elems[3 * i + k] = g.sm.tris[i][k];
			if (grad)
				grad[3 * i] = g.sm.normal[i].x, grad[3 * i + 1] = g.sm.normal[i].y, grad[3 * i + 2] = g.sm.normal[i].z;
		}
	}
	return 0;
}

int orc_set_pairs(void *h, const int *g1, const int *g2, int n)
{
	Scene &sc = *(Scene *)h;
	sc.pairs.clear();
	for (int i = 0; i < n; ++i) {
		if (g1[i] < 0 || g2[i] < 0 || g1[i] >= (int)sc.geoms.size() || g2[i] >= (int)sc.geoms.size())
			return -1;
		sc.pairs.push_back({ g1[i], g2[i] });
	}
	return 0;
}

// flat_tactile_sensor.cpp:127-214; geom_size = MuJoCo half sizes of the sensor box geom
int orc_add_flat_sensor(void *h, int geom, const double *geom_size, double resolution, int sampling_resolution,
                        int window, float sigma)
{
	Scene &sc = *(Scene *)h;
	FlatSensor fs;
	fs.geom       = geom;
	fs.resolution = resolution;
	fs.S          = sampling_resolution;
	fs.window     = window;
	fs.sigma      = sigma;
	if (window == 1 && sigma == -1.0f)
		fs.sigma = 0.1;
	if (window == 2 && sigma == -1.0f)
		fs.sigma = 0.3;
	for (int i = 0; i < 3; ++i)
		fs.size[i] = geom_size[i];
	fs.cx = (int)::floorl(2 * geom_size[0] / resolution + 0.1);
	fs.cy = (int)::floorl(2 * geom_size[1] / resolution + 0.1);
	sc.sensors.push_back(fs);
	return (int)sc.sensors.size() - 1;
}
int orc_sensor_dims(void *h, int sensor, int *cxcy)
{
	Scene &sc = *(Scene *)h;
	cxcy[0]   = sc.sensors[sensor].cx;
	cxcy[1]   = sc.sensors[sensor].cy;
	return 0;
}

// xpos[ng*3], xmat[ng*9] (row-major), vel[ng*6] = (omega, v) of mj_objectVelocity(..., flg_local=0)
int orc_step(void *h, const double *xpos, const double *xmat, const double *vel, int use_bvh)
{
	Scene &sc = *(Scene *)h;
	try {
		step(sc, sc.last, xpos, xmat, vel, use_bvh != 0);
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

// out[16] = {has_surface, gM, gN, n_faces, n_polygons, n_point_collisions, n_candidates,
//            F[3], tau[3] (about world origin, acting on gM), centroid[3], area}
int orc_pair_result(void *h, int pair, double *out)
{
	Scene &sc         = *(Scene *)h;
	const PairOut &po = sc.last.out[pair];
	std::memset(out, 0, 17 * sizeof(double));
	out[0] = po.has_surface;
	out[1] = po.gM;
	out[2] = po.gN;
	if (!po.has_surface)
		return 0;
	out[3]  = po.s->num_faces();
	out[4]  = (double)po.s->emitted.size();
	out[5]  = (double)po.pcs.size();
	out[6]  = (double)po.s->n_candidates;
	out[7]  = po.F.x, out[8] = po.F.y, out[9] = po.F.z;
	out[10] = po.tau.x, out[11] = po.tau.y, out[12] = po.tau.z;
	out[13] = po.centroid.x, out[14] = po.centroid.y, out[15] = po.centroid.z;
	out[16] = po.area;
	return 0;
}
// triples (elemM, elemN, nverts); returns count
int orc_pair_emitted(void *h, int pair, int *buf, int cap)
{
	Scene &sc         = *(Scene *)h;
	const PairOut &po = sc.last.out[pair];
	if (!po.has_surface)
		return 0;
	int n = (int)po.s->emitted.size();
	for (int i = 0; i < n && i < cap; ++i) {
		buf[3 * i]     = po.s->emitted[i].elemM;
		buf[3 * i + 1] = po.s->emitted[i].elemN;
		buf[3 * i + 2] = po.s->emitted[i].nverts;
	}
	return n;
}
// per PointCollision 13 doubles: p[3], n[3], fn0, stiffness, damping, face, f[3]; returns count
int orc_pair_faces(void *h, int pair, double *buf, int cap)
{
	Scene &sc         = *(Scene *)h;
	const PairOut &po = sc.last.out[pair];
	int n             = (int)po.pcs.size();
	for (int i = 0; i < n && i < cap; ++i) {
		const PointCollision &pc = po.pcs[i];
		double *o                = buf + 13 * i;
		o[0] = pc.p.x, o[1] = pc.p.y, o[2] = pc.p.z, o[3] = pc.n.x, o[4] = pc.n.y, o[5] = pc.n.z;
		o[6] = pc.fn0, o[7] = pc.stiffness, o[8] = pc.damping, o[9] = pc.face;
		o[10] = po.face_force[i].x, o[11] = po.face_force[i].y, o[12] = po.face_force[i].z;
	}
	return n;
}
// kTriangle surface soup: per face 9 doubles world vertices + 3 pressures; returns count
int orc_pair_triangles(void *h, int pair, double *buf, int cap)
{
	Scene &sc         = *(Scene *)h;
	const PairOut &po = sc.last.out[pair];
	if (!po.has_surface || !po.s->tri)
		return 0;
	const Surface &s = *po.s;
	int n            = s.num_faces();
	for (int i = 0; i < n && i < cap; ++i) {
		const int *f = &s.face_idx[s.face_first[i]];
		for (int k = 0; k < 3; ++k) {
			buf[12 * i + 3 * k]     = s.v[f[k]].x;
			buf[12 * i + 3 * k + 1] = s.v[f[k]].y;
			buf[12 * i + 3 * k + 2] = s.v[f[k]].z;
			buf[12 * i + 9 + k]     = s.e[f[k]];
		}
	}
	return n;
}
// The faces visualizeMeshElement(pc.face, mesh_W, fn) outlines (plugin.cpp:509-516, 525-555): per PointCollision
// 25 doubles = vertex count + up to 8 world vertices of face pc.face of the contact surface, in the face's winding.
int orc_pair_face_vertices(void *h, int pair, double *buf, int cap)
{
	Scene &sc         = *(Scene *)h;
	const PairOut &po = sc.last.out[pair];
	int n             = (int)po.pcs.size();
	for (int i = 0; i < n && i < cap; ++i) {
		const Surface &s = *po.s;
		int face = po.pcs[i].face, nv = s.face_n[face];
		const int *f = &s.face_idx[s.face_first[face]];
		double *o    = buf + 25 * i;
		for (int k = 0; k < 25; ++k)
			o[k] = 0;
		o[0] = nv;
		for (int k = 0; k < nv && k < 8; ++k)
			o[1 + 3 * k] = s.v[f[k]].x, o[2 + 3 * k] = s.v[f[k]].y, o[3 + 3 * k] = s.v[f[k]].z;
	}
	return n;
}
int orc_geom_wrench(void *h, int geom, double *out6)
{
	Scene &sc = *(Scene *)h;
	for (int i = 0; i < 6; ++i)
		out6[i] = sc.last.geom_wrench[geom][i];
	return 0;
}
// curved_sensor.cpp:111-380 (load, with the surface samples supplied by the caller) and :388-481
int orc_add_curved_sensor(void *h, int geom, int n_taxels, const double *taxel_pos, const double *taxel_nrm,
                          int n_samples, const double *sample_pos, const double *sample_nrm, double include_margin)
{
	Scene &sc = *(Scene *)h;
	CurvedSensor cs;
	cs.geom           = geom;
	cs.include_margin = include_margin;
	for (int i = 0; i < n_taxels; ++i) {
		cs.taxel_pos.push_back({ taxel_pos[3 * i], taxel_pos[3 * i + 1], taxel_pos[3 * i + 2] });
		V3 nn{ 0, 0, 0 };
		if (taxel_nrm) { // normalised like :187
			nn = { taxel_nrm[3 * i], taxel_nrm[3 * i + 1], taxel_nrm[3 * i + 2] };
			double l = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
			if (l > 0)
				nn = { nn[0] / l, nn[1] / l, nn[2] / l };
		}
		cs.taxel_nrm.push_back(nn);
	}
	curved_sensor_load(cs, sample_pos, sample_nrm, n_samples);
	sc.curved.push_back(cs);
	return (int)sc.curved.size() - 1;
}
int orc_curved_values(void *h, int sensor, float *out, int use_bvh)
{
	Scene &sc = *(Scene *)h;
	curved_sensor_values(sc, sc.last, sensor, out, use_bvh);
	return 0;
}
int orc_curved_info(void *h, int sensor, int *n_close_n_assign)
{
	Scene &sc = *(Scene *)h;
	const CurvedSensor &cs = sc.curved[sensor];
	size_t a = 0;
	for (const auto &v : cs.surface_idx)
		a += v.size();
	n_close_n_assign[0] = (int)cs.surf_pos.size();
	n_close_n_assign[1] = (int)a;
	return 0;
}

// taxel_sensor.cpp:45-156 (load) and :158-478
int orc_add_taxel_sensor(void *h, int geom, int n_taxels, const double *taxel_pos, double include_margin,
                         double sample_resolution, int method, int visualize, int sample_method)
{
	Scene &sc = *(Scene *)h;
	TaxelSensor ts;
	ts.geom = geom, ts.include_margin = include_margin, ts.sample_resolution = sample_resolution;
	ts.method = method, ts.visualize = visualize != 0, ts.sample_method = sample_method;
	for (int i = 0; i < n_taxels; ++i)
		ts.taxels.push_back({ taxel_pos[3 * i], taxel_pos[3 * i + 1], taxel_pos[3 * i + 2] });
	sc.taxel.push_back(ts);
	return (int)sc.taxel.size() - 1;
}
int orc_taxel_values(void *h, int sensor, float *values_inout)
{
	Scene &sc = *(Scene *)h;
	taxel_sensor_values(sc, sc.last, sensor, values_inout);
	return 0;
}

// use_bvh: 0 linear scan, 1 the oracle's BVH/TLAS restatement, 2 the reference's compiled ray caster (oracle/_ref,
// installed with orc_set_external_caster)
int orc_sensor_image(void *h, int sensor, float *out, int use_bvh, int parallel)
{
	Scene &sc = *(Scene *)h;
	try {
		flat_sensor_image(sc, sc.last, sensor, out, use_bvh, parallel != 0);
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

// as orc_sensor_image, plus every ray and its nearest hit: rays[n][6] = (O, D), tuv[n][3], id[n], n = cx*cy*S*S in
// (x, y, i, j) order
int orc_sensor_image_trace(void *h, int sensor, float *out, int use_bvh, float *rays, float *tuv, uint32_t *id)
{
	Scene &sc = *(Scene *)h;
	try {
		FlatTrace tr;
		flat_sensor_image(sc, sc.last, sensor, out, use_bvh, false, &tr);
		if (tr.id.empty()) { // no surface touches the sensor
			const FlatSensor &fs = sc.sensors[sensor];
			size_t n             = (size_t)fs.cx * fs.cy * fs.S * fs.S;
			std::fill(rays, rays + 6 * n, 0.0f);
			for (size_t i = 0; i < n; ++i)
				tuv[3 * i] = 1e30f, tuv[3 * i + 1] = tuv[3 * i + 2] = 0, id[i] = 0;
			return 0;
		}
		std::copy(tr.rays.begin(), tr.rays.end(), rays);
		std::copy(tr.tuv.begin(), tr.tuv.end(), tuv);
		std::copy(tr.id.begin(), tr.id.end(), id);
	} catch (const std::exception &e) {
		g_err = e.what();
		return -1;
	}
	return 0;
}

// the three entry points of oracle/_ref/libref_bvh*.so (ref_tlas_create / ref_tlas_cast / ref_tlas_destroy)
int orc_set_external_caster(void *create, void *cast, void *destroy)
{
	ExternalCaster c;
	c.create  = (decltype(c.create))create;
	c.cast    = (decltype(c.cast))cast;
	c.destroy = (decltype(c.destroy))destroy;
	set_external_caster(c);
	return 0;
}

int orc_cast_rays(int n_surf, const int *n_tri, const double *verts, int n_rays, const float *O, const float *D,
                  float *tuv, uint32_t *id)
{
	cast_rays(n_surf, n_tri, verts, n_rays, O, D, tuv, id);
	return 0;
}
void orc_intersect_triangle(const float *O, const float *D, const float *v0, const float *v1, const float *v2, float t_in,
                            float *tuv_out, int *hit_out)
{
	intersect_triangle_one(O, D, v0, v1, v2, t_in, tuv_out, hit_out);
}
float orc_intersect_aabb(const float *O, const float *D, float t_in, const float *bmin, const float *bmax)
{
	return intersect_aabb_one(O, D, t_in, bmin, bmax);
}

// CPU baseline: n_env independent env steps (plus every sensor image when with_sensors), envs
// distributed over `threads` OpenMP threads (1 = the reference's execution model: one physics
// thread).  out = {seconds, total candidate pair-evals, checksum of |F|}
int orc_bench(void *h, int n_env, const double *xpos, const double *xmat, const double *vel, int use_bvh,
              int with_sensors, int threads, double *out)
{
	Scene &sc = *(Scene *)h;
	int ng    = (int)sc.geoms.size();
	long cands = 0;
	double checksum = 0;
	auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
	if (threads < 1)
		threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads) reduction(+ : cands, checksum)
#endif
	{
		StepState st;
		std::vector<float> img;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
		for (int e = 0; e < n_env; ++e) {
			step(sc, st, xpos + (size_t)e * ng * 3, xmat + (size_t)e * ng * 9, vel + (size_t)e * ng * 6, use_bvh != 0);
			cands += st.n_candidates;
			for (auto &po : st.out)
				checksum += norm(po.F);
			if (with_sensors)
				for (size_t si = 0; si < sc.sensors.size(); ++si) {
					img.resize(sc.sensors[si].cx * sc.sensors[si].cy);
					// inside an env-parallel region the taxel loop runs serially
					flat_sensor_image(sc, st, (int)si, img.data(), true, threads == 1);
					checksum += img[img.size() / 2];
				}
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	out[0]  = std::chrono::duration<double>(t1 - t0).count();
	out[1]  = (double)cands;
	out[2]  = checksum;
	return 0;
}

int orc_num_threads()
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} // extern "C"
