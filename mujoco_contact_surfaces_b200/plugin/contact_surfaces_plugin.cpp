// Host adapter: the reference's MujocoContactSurfacesPlugin / SurfacePlugin / FlatTactileSensor surface on
// top of the CUDA engine (libhcs_b200, include/hcs.h).  See contact_surfaces_plugin.h for the mapping.
//
// Differences to the reference's control flow (behaviour preserved):
//  * collision_cb (plugin.cpp:255-318) only RECORDS the geom pair MuJoCo hands over and returns the same
//    contact count (0 when forces are applied); the contact surfaces of all recorded pairs are computed in
//    ONE batched GPU call at the start of passiveCallback, where the reference consumes them (:411-523).
//  * two mj_applyFT per face (:477-482) become one mj_applyFT per geom with the reduced wrench — exact,
//    because mj_applyFT is linear in (force, torque about the application point).
//  * Q1 (numeric_size indexed by address, :596,:606) and Q2 (onGeomChanged drops the new pressure field,
//    :842-845) are consciously fixed, not reproduced.
#include "contact_surfaces_plugin.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <unordered_map>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mujoco_ros::contact_surfaces {

namespace {
// plugin.cpp:88-95
std::map<const mjData *, MujocoContactSurfacesPlugin *> instance_map;
mjfCollision defaultCollisionFunctions[mjNGEOMTYPES][mjNGEOMTYPES];

int collision_cb_wrapper(const mjModel *m, const mjData *d, mjContact *con, int g1, int g2, mjtNum margin)
{
	return instance_map[d]->collision_cb(m, d, con, g1, g2, margin);
}

void log_info(const char *fmt, const char *a = "", const char *b = "")
{
	if (std::getenv("HCS_PLUGIN_VERBOSE")) {
		std::fprintf(stderr, "[mujoco_contact_surfaces] ");
		std::fprintf(stderr, fmt, a, b);
		std::fprintf(stderr, "\n");
	}
}
} // namespace

MujocoContactSurfacesPlugin::MujocoContactSurfacesPlugin() {}

// plugin.cpp:191-206
MujocoContactSurfacesPlugin::~MujocoContactSurfacesPlugin()
{
	geomCollisions.clear();
	contactProperties.clear();
	instance_map.erase(d_);
	if (instance_map.empty()) {
		for (int i = 0; i < mjNGEOMTYPES; ++i)
			for (int j = 0; j < mjNGEOMTYPES; ++j)
				if (mjCOLLISIONFUNC[i][j] == collision_cb_wrapper)
					mjCOLLISIONFUNC[i][j] = defaultCollisionFunctions[i][j];
	}
	if (ctx_)
		hcs_destroy(ctx_);
	delete[] vGeoms;
}

void MujocoContactSurfacesPlugin::addSurfacePlugin(SurfacePluginPtr plugin, const PluginConfig &config)
{
	plugin->init(config, "");
	plugin->attach(this);
	plugins.push_back(plugin);
}

int MujocoContactSurfacesPlugin::configIndex(int mujoco_geom_id) const
{
	auto it = contactProperties.find(mujoco_geom_id);
	return it == contactProperties.end() ? -1 : it->second->drake_id;
}

// plugin.cpp:208-245
bool MujocoContactSurfacesPlugin::load(const mjModel *m, mjData *d)
{
	parseMujocoCustomFields(m);
	if (!ctx_)
		return false; // no CUDA device: there is no CPU fallback
	d_              = d;
	m_              = m;
	instance_map[d] = this;
	if (instance_map.size() == 1)
		initCollisionFunction();
	for (const auto &plugin : plugins)
		if (plugin->safe_load(m, d))
			cb_ready_plugins.push_back(plugin);
	return true;
}

// plugin.cpp:247-253
void MujocoContactSurfacesPlugin::reset()
{
	geomCollisions.clear();
	step_pairs_.clear();
	for (const auto &plugin : plugins)
		plugin->safe_reset();
}

// plugin.cpp:557-569
void MujocoContactSurfacesPlugin::initCollisionFunction()
{
	for (int i = 0; i < mjNGEOMTYPES; ++i)
		for (int j = 0; j < mjNGEOMTYPES; ++j) {
			defaultCollisionFunctions[i][j] = mjCOLLISIONFUNC[i][j];
			mjCOLLISIONFUNC[i][j]           = collision_cb_wrapper; // env_ptr_->registerCollisionFunction(i, j, ...)
		}
}

// plugin.cpp:571-813
void MujocoContactSurfacesPlugin::parseMujocoCustomFields(const mjModel *m)
{
	int hcp_id = mj_name2id(m, mjOBJ_TEXT, (PREFIX + "HydroelasticContactRepresentation").c_str());
	if (hcp_id >= 0 && m->text_adr[hcp_id] >= 0) {
		std::string hcp(&m->text_data[m->text_adr[hcp_id]], m->text_size[hcp_id]);
		if (hcp.find("kTriangle") != std::string::npos)
			hydroelastic_contact_representation = HCS_REP_TRIANGLE;
		else if (hcp.find("kPolygon") != std::string::npos)
			hydroelastic_contact_representation = HCS_REP_POLYGON;
	}
	int vs_id = mj_name2id(m, mjOBJ_NUMERIC, (PREFIX + "VisualizeSurfaces").c_str());
	if (vs_id >= 0 && m->numeric_size[vs_id] == 1)
		visualizeContactSurfaces = m->numeric_data[m->numeric_adr[vs_id]] != 0;
	int apsf_id = mj_name2id(m, mjOBJ_NUMERIC, (PREFIX + "ApplyContactSurfaceForces").c_str());
	if (apsf_id >= 0 && m->numeric_size[apsf_id] == 1)
		applyContactSurfaceForces = m->numeric_data[m->numeric_adr[apsf_id]] != 0;

	hcs_config cfg;
	std::memset(&cfg, 0, sizeof cfg);
	cfg.device               = device;
	cfg.n_envs               = 1; // one plugin instance serves one mjData (plugin.cpp:88, 225)
	cfg.representation       = hydroelastic_contact_representation;
	cfg.apply_contact_forces = applyContactSurfaceForces;
	cfg.max_faces            = 1 << 16; // GeomCollision views for sub-plugins / visualisation
	cfg.face_vertices        = visualizeContactSurfaces ? 1 : 0;
	if (hcs_create(&cfg, &ctx_) != HCS_OK) {
		std::fprintf(stderr, "[mujoco_contact_surfaces] %s\n", hcs_last_error(nullptr));
		ctx_ = nullptr;
		return;
	}
	// per-geom contact properties, in the order of the numerics = drake_id order
	for (int i = 0; i < m->nnumeric; ++i) {
		const char *nm = mj_id2name(m, mjOBJ_NUMERIC, i);
		if (!nm)
			continue;
		std::string full_name = nm;
		if (full_name.rfind(PREFIX, 0) != 0)
			continue;
		std::string s = full_name.substr(PREFIX.length());
		int id        = mj_name2id(m, mjOBJ_GEOM, s.c_str());
		if (id < 0)
			continue;
		int adr = m->numeric_adr[i], size = m->numeric_size[i];
		if (adr < 0 || size != 5)
			continue;
		const double *props = m->numeric_data + adr;
		const float *mv     = nullptr;
		const int32_t *mf   = nullptr;
		int nv = 0, nf = 0;
		if (m->geom_type[id] == mjGEOM_MESH) {
			int did = m->geom_dataid[id];
			if (did >= 0) {
				nv = m->mesh_vertnum[did];
				nf = m->mesh_facenum[did];
				mv = m->mesh_vert + 3 * m->mesh_vertadr[did];
				mf = m->mesh_face + 3 * m->mesh_faceadr[did];
			}
		}
		int cfg_idx = hcs_add_geom(ctx_, m->geom_type[id], m->geom_size + 3 * id, mv, nv, mf, nf, props);
		if (cfg_idx < 0) { // plane-soft / hfield / capsule / bad mesh: skip, MuJoCo's default collision stays
			log_info("geom '%s' skipped: %s", s.c_str(), hcs_last_error(ctx_));
			continue;
		}
		auto cp             = std::make_shared<ContactProperties>();
		cp->mujoco_geom_id  = id;
		cp->drake_id        = cfg_idx;
		cp->geom_name       = s;
		cp->contact_type    = props[0] > 0 ? SOFT : RIGID;
		cp->hydroelastic_modulus = props[0] > 0 ? props[0] : INFINITY;
		cp->dissipation     = props[0] > 0 ? props[1] : 1.0;
		cp->resolution_hint = props[2];
		cp->static_friction = props[3];
		cp->dynamic_friction = props[4];
		contactProperties[id] = cp;
		cfg_to_mj.push_back(id);
	}
}

// plugin.cpp:255-318: only the dispatch decision stays on the host
int MujocoContactSurfacesPlugin::collision_cb(const mjModel *m, const mjData *d, mjContact *con, int g1, int g2,
                                              mjtNum margin)
{
	int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
	auto c1 = contactProperties.find(g1), c2 = contactProperties.find(g2);
	if (c1 == contactProperties.end() || c2 == contactProperties.end() ||
	    (c1->second->contact_type == RIGID && c2->second->contact_type == RIGID))
		return defaultCollisionFunctions[t1][t2](m, d, con, g1, g2, margin);
	int n_con = applyContactSurfaceForces ? 0 : defaultCollisionFunctions[t1][t2](m, d, con, g1, g2, margin);
	step_pairs_.emplace_back(c1->second->drake_id, c2->second->drake_id);
	return n_con;
}

void MujocoContactSurfacesPlugin::ensurePairs()
{
	bool changed = !finalized_;
	for (const auto &p : step_pairs_)
		if (known_pairs_.insert(p).second) {
			pair_list_.push_back(p);
			changed = true;
		}
	if (!changed)
		return;
	std::vector<int32_t> g1, g2;
	for (const auto &p : pair_list_) {
		g1.push_back(p.first);
		g2.push_back(p.second);
	}
	if (hcs_set_pairs(ctx_, g1.data(), g2.data(), (int)g1.size()) != HCS_OK || hcs_finalize(ctx_) != HCS_OK) {
		std::fprintf(stderr, "[mujoco_contact_surfaces] %s\n", hcs_last_error(ctx_));
		finalized_ = false;
		return;
	}
	finalized_ = true;
}

void MujocoContactSurfacesPlugin::evaluateAndApply(const mjModel *m, mjData *d, bool with_sensors)
{
	int ng = (int)cfg_to_mj.size();
	xpos_.resize(3 * ng), xmat_.resize(9 * ng), vel_.resize(6 * ng), wrench_.resize(6 * ng);
	for (int c = 0; c < ng; ++c) {
		int id = cfg_to_mj[c];
		std::memcpy(&xpos_[3 * c], d->geom_xpos + 3 * id, 3 * sizeof(double)); // getGeomPose, plugin.cpp:116-126
		std::memcpy(&xmat_[9 * c], d->geom_xmat + 9 * id, 9 * sizeof(double));
		mj_objectVelocity(m, d, mjOBJ_GEOM, id, &vel_[6 * c], 0); // getGeomVelocity, plugin.cpp:107-114
	}
	if (hcs_step(ctx_, xpos_.data(), xmat_.data(), vel_.data(), with_sensors ? 1 : 0) != HCS_OK) {
		std::fprintf(stderr, "[mujoco_contact_surfaces] %s\n", hcs_last_error(ctx_));
		return;
	}
	if (applyContactSurfaceForces) { // plugin.cpp:477-482, reduced to one wrench per geom
		hcs_get_geom_wrenches(ctx_, wrench_.data());
		const mjtNum origin[3] = { 0, 0, 0 };
		for (int c = 0; c < ng; ++c) {
			const double *w = &wrench_[6 * c];
			if (w[0] == 0 && w[1] == 0 && w[2] == 0 && w[3] == 0 && w[4] == 0 && w[5] == 0)
				continue;
			mj_applyFT(m, d, w, w + 3, origin, m->geom_bodyid[cfg_to_mj[c]], d->qfrc_passive);
		}
	}
}

// GeomCollision / PointCollision views (common_types.h:48-66) from the per-face dump
void MujocoContactSurfacesPlugin::buildGeomCollisions()
{
	geomCollisions.clear();
	int np = hcs_n_pairs(ctx_);
	if (np <= 0)
		return;
	pair_results_.resize(np);
	if (hcs_get_pair_results(ctx_, pair_results_.data()) != HCS_OK)
		return;
	if (faces_.empty())
		faces_.resize(1 << 16); // allocated once (cfg.max_faces); the dump only overwrites what it returns
	std::vector<hcs_face> &faces = faces_;
	int nf = hcs_get_faces(ctx_, faces.data(), (int)faces.size());
	if (nf < 0)
		nf = 0;
	nf = std::min<int>(nf, (int)faces.size());
	std::vector<double> fverts; // the same faces' world vertices (only kept when the surfaces are drawn)
	if (visualizeContactSurfaces && nf > 0) {
		fverts.resize((size_t)nf * HCS_FACE_VERTEX_STRIDE);
		if (hcs_get_face_vertices(ctx_, fverts.data(), nf) < 0)
			fverts.clear();
	}
	std::vector<int> order(nf); // the dump's order is unspecified: (pair, elemM, elemN, face) is Drake's face order
	for (int i = 0; i < nf; ++i)
		order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&faces](int ia, int ib) {
		const hcs_face &a = faces[ia], &b = faces[ib];
		if (a.pair != b.pair) return a.pair < b.pair;
		if (a.elemM != b.elemM) return a.elemM < b.elemM;
		if (a.elemN != b.elemN) return a.elemN < b.elemN;
		return a.face < b.face;
	});
	std::vector<double> tris;
	std::vector<int32_t> tri_pair; // the soup is sorted by pair: every surface gets its own run
	int ntri = 0;
	if (hydroelastic_contact_representation == HCS_REP_TRIANGLE) {
		ntri = hcs_get_tactile_triangles(ctx_, 0, nullptr, 0);
		if (ntri > 0) {
			tris.resize(12 * (size_t)ntri);
			tri_pair.resize(ntri);
			hcs_get_tactile_triangles(ctx_, 0, tris.data(), ntri);
			hcs_get_tactile_triangle_pairs(ctx_, 0, tri_pair.data(), ntri);
		}
	}
	for (int p = 0; p < np; ++p) {
		const hcs_pair_result &r = pair_results_[p];
		if (r.n_polygons == 0)
			continue; // the reference only stores non-null surfaces (plugin.cpp:307-315)
		auto *view        = new ContactSurfaceView();
		view->is_triangle = hydroelastic_contact_representation == HCS_REP_TRIANGLE;
		view->total_area  = r.area;
		view->num_faces   = r.n_faces;
		std::memcpy(view->centroid, r.centroid, sizeof view->centroid);
		for (int t = 0; t < ntri; ++t) // this surface's triangles (12 doubles each: vertices + pressures)
			if (tri_pair[t] == p)
				view->triangles.insert(view->triangles.end(), tris.begin() + 12 * (size_t)t, tris.begin() + 12 * (size_t)(t + 1));
		GeomCollisionPtr gc(new GeomCollision(cfg_to_mj[r.gM], cfg_to_mj[r.gN], view));
		int k = 0;
		for (int j = 0; j < nf; ++j) {
			const hcs_face &f = faces[order[j]];
			if (f.pair != p)
				continue;
			PointCollision pc;
			std::memcpy(pc.p, f.p, sizeof pc.p);
			std::memcpy(pc.n, f.n, sizeof pc.n);
			pc.fn0 = f.fn0, pc.stiffness = f.stiffness, pc.damping = f.damping;
			pc.face = k++;
			gc->pointCollisions.push_back(pc);
			// fn of plugin.cpp:470 (the friction part of f is tangential)
			view->face_fn.push_back(f.f[0] * f.n[0] + f.f[1] * f.n[1] + f.f[2] * f.n[2]);
			if (!fverts.empty()) {
				view->face_nverts.push_back(view->is_triangle ? 3 : f.nverts);
				const double *v = &fverts[(size_t)order[j] * HCS_FACE_VERTEX_STRIDE];
				view->face_vertices.insert(view->face_vertices.end(), v, v + HCS_FACE_VERTEX_STRIDE);
			}
		}
		geomCollisions.push_back(gc);
	}
}

// plugin.cpp:525-555: the outline of one face of the contact surface, one thin cylinder per edge
void MujocoContactSurfacesPlugin::visualizeMeshElement(int face, const ContactSurfaceView &mesh, double /*fn*/)
{
	if (face >= (int)mesh.face_nverts.size())
		return;
	if (n_vGeom >= MAX_VGEOM) {
		std::fprintf(stderr, "n_vGeom too big\n");
		return;
	}
	const float rgba[4] = { 0.3f, 0.3f, 0.3f, 0.8f };
	const int nv        = mesh.face_nverts[face];
	const double *v     = &mesh.face_vertices[(size_t)face * HCS_FACE_VERTEX_STRIDE];
	const double *vp0   = v + 3 * (nv - 1);
	for (int i = 0; i < nv; ++i) {
		const double *vp1 = v + 3 * i;
		if (n_vGeom == MAX_VGEOM) {
			std::fprintf(stderr, "n_vGeom too big\n");
			break;
		}
		mjvGeom *g = vGeoms + n_vGeom++;
		mjv_initGeom(g, mjGEOM_CYLINDER, nullptr, nullptr, nullptr, rgba);
		mjv_makeConnector(g, mjGEOM_CYLINDER, 0.000015, vp0[0], vp0[1], vp0[2], vp1[0], vp1[1], vp1[2]);
		vp0 = vp1;
	}
}

// plugin.cpp:411-523
void MujocoContactSurfacesPlugin::passiveCallback(const mjModel *m, mjData *d)
{
	if (!ctx_)
		return;
	if (visualizeContactSurfaces) {
		n_vGeom       = 0;
		running_scale = 0.9 * running_scale + 0.1 * current_scale;
		current_scale = 0.;
	}
	ensurePairs();
	bool with_sensors = false;
	for (const auto &plugin : cb_ready_plugins)
		with_sensors |= plugin->wantsSensorUpdate(d);
	if (finalized_ && !pair_list_.empty()) {
		evaluateAndApply(m, d, with_sensors);
		// GeomCollision views exist for their consumers (CPU sub-plugins, the surface outlines): nobody to read them,
		// nothing to fetch (the forces have been applied from the per-geom wrenches already)
		if (visualizeContactSurfaces || !cb_ready_plugins.empty())
			buildGeomCollisions();
		if (visualizeContactSurfaces) { // plugin.cpp:509-516
			for (const auto &gc : geomCollisions)
				for (const auto &pc : gc->pointCollisions) {
					const double fn = gc->s->face_fn[pc.face];
					current_scale   = std::max(current_scale, std::fabs(fn));
					visualizeMeshElement(pc.face, *gc->s, fn);
				}
		}
	}
	for (const auto &plugin : cb_ready_plugins)
		plugin->update(m, d, geomCollisions);
	geomCollisions.clear();
	step_pairs_.clear();
}

// plugin.cpp:815-826
void MujocoContactSurfacesPlugin::renderCallback(const mjModel *model, mjData *data, mjvScene *scene)
{
	if (visualizeContactSurfaces) {
		int n = std::min(n_vGeom, scene->maxgeom - scene->ngeom);
		for (int i = 0; i < n; ++i)
			scene->geoms[scene->ngeom++] = vGeoms[i];
	}
	for (const auto &plugin : cb_ready_plugins)
		plugin->renderCallback(model, data, scene);
}

// plugin.cpp:828-975 (the new mesh AND its pressure field are installed: Q2 fixed)
void MujocoContactSurfacesPlugin::onGeomChanged(const mjModel *m, mjData *, const int id)
{
	auto it = contactProperties.find(id);
	if (it == contactProperties.end() || !ctx_)
		return;
	if (hcs_update_geom(ctx_, it->second->drake_id, m->geom_size + 3 * id) != HCS_OK)
		std::fprintf(stderr, "[mujoco_contact_surfaces] %s\n", hcs_last_error(ctx_));
}

// ---- batched host adapter -------------------------------------------------------------------------------------
BatchedContactSurfaces::~BatchedContactSurfaces()
{
	if (ctx_)
		hcs_multi_destroy(ctx_);
}

// the custom-field walk of parseMujocoCustomFields (plugin.cpp:571-813) for a multi-device context
bool BatchedContactSurfaces::load(const mjModel *m, int n_envs, const std::vector<int> &devices)
{
	int representation = HCS_REP_TRIANGLE;
	int hcp_id         = mj_name2id(m, mjOBJ_TEXT, (PREFIX + "HydroelasticContactRepresentation").c_str());
	if (hcp_id >= 0 && m->text_adr[hcp_id] >= 0) {
		std::string hcp(&m->text_data[m->text_adr[hcp_id]], m->text_size[hcp_id]);
		if (hcp.find("kPolygon") != std::string::npos)
			representation = HCS_REP_POLYGON;
	}
	int apsf_id = mj_name2id(m, mjOBJ_NUMERIC, (PREFIX + "ApplyContactSurfaceForces").c_str());
	if (apsf_id >= 0 && m->numeric_size[apsf_id] == 1)
		applyContactSurfaceForces = m->numeric_data[m->numeric_adr[apsf_id]] != 0;
	hcs_config cfg;
	std::memset(&cfg, 0, sizeof cfg);
	cfg.n_envs               = n_envs;
	cfg.representation       = representation;
	cfg.apply_contact_forces = applyContactSurfaceForces;
	if (hcs_multi_create(&cfg, devices.data(), (int)devices.size(), &ctx_) != HCS_OK) {
		std::fprintf(stderr, "[mujoco_contact_surfaces] %s\n", hcs_multi_last_error(nullptr));
		ctx_ = nullptr;
		return false;
	}
	n_envs_ = n_envs;
	for (int i = 0; i < m->nnumeric; ++i) {
		const char *nm = mj_id2name(m, mjOBJ_NUMERIC, i);
		if (!nm)
			continue;
		std::string full_name = nm;
		if (full_name.rfind(PREFIX, 0) != 0)
			continue;
		std::string s = full_name.substr(PREFIX.length());
		int id        = mj_name2id(m, mjOBJ_GEOM, s.c_str());
		int adr = m->numeric_adr[i], size = m->numeric_size[i];
		if (id < 0 || adr < 0 || size != 5)
			continue;
		const double *props = m->numeric_data + adr;
		const float *mv     = nullptr;
		const int32_t *mf   = nullptr;
		int nv = 0, nf = 0;
		if (m->geom_type[id] == mjGEOM_MESH && m->geom_dataid[id] >= 0) {
			int did = m->geom_dataid[id];
			nv = m->mesh_vertnum[did], nf = m->mesh_facenum[did];
			mv = m->mesh_vert + 3 * m->mesh_vertadr[did];
			mf = m->mesh_face + 3 * m->mesh_faceadr[did];
		}
		int cfg_idx = hcs_multi_add_geom(ctx_, m->geom_type[id], m->geom_size + 3 * id, mv, nv, mf, nf, props);
		if (cfg_idx < 0)
			continue; // unsupported geom: MuJoCo's default collision stays
		auto cp              = std::make_shared<ContactProperties>();
		cp->mujoco_geom_id   = id;
		cp->drake_id         = cfg_idx;
		cp->geom_name        = s;
		cp->contact_type     = props[0] > 0 ? SOFT : RIGID;
		cp->hydroelastic_modulus = props[0] > 0 ? props[0] : INFINITY;
		cp->dissipation      = props[0] > 0 ? props[1] : 1.0;
		cp->resolution_hint  = props[2];
		cp->static_friction  = props[3];
		cp->dynamic_friction = props[4];
		contactProperties[id] = cp;
		cfg_to_mj.push_back(id);
	}
	return true;
}

void BatchedContactSurfaces::recordPair(int g1, int g2)
{
	auto c1 = contactProperties.find(g1), c2 = contactProperties.find(g2);
	if (c1 == contactProperties.end() || c2 == contactProperties.end() ||
	    (c1->second->contact_type == RIGID && c2->second->contact_type == RIGID))
		return;
	std::pair<int, int> p(c1->second->drake_id, c2->second->drake_id);
	if (known_pairs_.insert(p).second) {
		pair_list_.push_back(p);
		finalized_ = false;
	}
}

bool BatchedContactSurfaces::finalize()
{
	if (!ctx_)
		return false;
	std::vector<int32_t> g1, g2;
	for (const auto &p : pair_list_) {
		g1.push_back(p.first);
		g2.push_back(p.second);
	}
	if (hcs_multi_set_pairs(ctx_, g1.data(), g2.data(), (int)g1.size()) != HCS_OK || hcs_multi_finalize(ctx_) != HCS_OK) {
		std::fprintf(stderr, "[mujoco_contact_surfaces] %s\n", hcs_multi_last_error(ctx_));
		return false;
	}
	finalized_ = true;
	return true;
}

void BatchedContactSurfaces::passiveCallback(const mjModel *m, mjData *const *d, int n)
{
	if (!ctx_ || n != n_envs_ || pair_list_.empty() || (!finalized_ && !finalize()))
		return;
	const int ng = (int)cfg_to_mj.size();
	xpos_.resize((size_t)3 * ng * n), xmat_.resize((size_t)9 * ng * n), vel_.resize((size_t)6 * ng * n), wrench_.resize((size_t)6 * ng * n);
	for (int e = 0; e < n; ++e)
		for (int c = 0; c < ng; ++c) {
			const int id = cfg_to_mj[c];
			const size_t o = (size_t)e * ng + c;
			std::memcpy(&xpos_[3 * o], d[e]->geom_xpos + 3 * id, 3 * sizeof(double));
			std::memcpy(&xmat_[9 * o], d[e]->geom_xmat + 9 * id, 9 * sizeof(double));
			mj_objectVelocity(m, d[e], mjOBJ_GEOM, id, &vel_[6 * o], 0);
		}
	if (hcs_multi_step(ctx_, xpos_.data(), xmat_.data(), vel_.data(), 0) != HCS_OK) {
		std::fprintf(stderr, "[mujoco_contact_surfaces] %s\n", hcs_multi_last_error(ctx_));
		return;
	}
	if (!applyContactSurfaceForces)
		return;
	hcs_multi_get_geom_wrenches(ctx_, wrench_.data());
	const mjtNum origin[3] = { 0, 0, 0 };
	for (int e = 0; e < n; ++e)
		for (int c = 0; c < ng; ++c) {
			const double *w = &wrench_[6 * ((size_t)e * ng + c)];
			if (w[0] == 0 && w[1] == 0 && w[2] == 0 && w[3] == 0 && w[4] == 0 && w[5] == 0)
				continue;
			mj_applyFT(m, d[e], w, w + 3, origin, m->geom_bodyid[cfg_to_mj[c]], d[e]->qfrc_passive);
		}
}

namespace sensors {

static bool has(const PluginConfig &c, const char *k) { return c.find(k) != c.end(); }

// tactile_sensor_base.cpp:61-87
bool TactileSensorBase::load(const mjModel *m, mjData *)
{
	const PluginConfig &c = rosparam_config_;
	if (!(has(c, "geomName") && has(c, "topicName") && has(c, "updateRate") && has(c, "sensorName")))
		return false;
	geomName = c.at("geomName");
	int id   = mj_name2id(m, mjOBJ_GEOM, geomName.c_str());
	if (id < 0)
		return false;
	geomID       = id;
	lastUpdate   = -1e300;
	topicName    = c.at("topicName");
	sensorName   = c.at("sensorName");
	updateRate   = std::atof(c.at("updateRate").c_str());
	updatePeriod = 1.0 / updateRate;
	if (has(c, "visualize"))
	{ // YAML booleans as the reference's configs write them (fingertip.yaml: "visualize: True")
		std::string v = c.at("visualize");
		for (char &ch : v)
			ch = (char)std::tolower((unsigned char)ch);
		visualize = v == "true" || v == "1" || v == "yes" || v == "on";
	}
	return true;
}

bool TactileSensorBase::wantsSensorUpdate(const mjData *d)
{
	std::lock_guard<std::mutex> pause_lock(pause_mutex);
	double last = d->time < lastUpdate ? -1e300 : lastUpdate;
	bool due    = (d->time - last >= updatePeriod) && !paused;
	std::lock_guard<std::mutex> state_lock(state_request_mutex);
	return due || request_state;
}

// tactile_sensor_base.cpp:89-120
void TactileSensorBase::update(const mjModel *m, mjData *d, const std::vector<GeomCollisionPtr> &geomCollisions)
{
	std::lock_guard<std::mutex> pause_lock(pause_mutex);
	double now = d->time;
	if (now < lastUpdate)
		lastUpdate = -1e300; // reset lastUpdate after jump back in time
	if (now - lastUpdate >= updatePeriod && !paused) {
		lastUpdate = now;
		n_vGeom    = 0;
		internal_update(m, d, geomCollisions);
		++publish_count_; // publisher.publish(tactile_state_msg_)
	}
	std::unique_lock<std::mutex> state_lock(state_request_mutex);
	if (request_state) {
		n_vGeom = 0;
		internal_update(m, d, geomCollisions);
		request_state = false;
		state_lock.unlock();
		state_cv.notify_one();
	}
}

void TactileSensorBase::renderCallback(const mjModel *, mjData *, mjvScene *scene)
{
	if (visualize)
		for (int i = 0; i < n_vGeom && scene->ngeom < scene->maxgeom; ++i)
			scene->geoms[scene->ngeom++] = vGeoms[i];
}

void TactileSensorBase::reset() {}

void TactileSensorBase::setPause(bool pause)
{
	std::lock_guard<std::mutex> pause_lock(pause_mutex);
	paused  = pause;
	n_vGeom = 0;
}

std::vector<float> TactileSensorBase::getState()
{
	std::unique_lock<std::mutex> state_lock(state_request_mutex);
	request_state = true;
	state_cv.wait(state_lock, [this] { return !request_state; });
	return tactile_state_values_;
}

// flat_tactile_sensor.cpp:127-214
bool FlatTactileSensor::load(const mjModel *m, mjData *d)
{
	const PluginConfig &c = rosparam_config_;
	if (!(TactileSensorBase::load(m, d) && has(c, "resolution")) || !owner_ || !owner_->context())
		return false;
	resolution = std::atof(c.at("resolution").c_str());
	if (has(c, "sampling_resolution"))
		sampling_resolution = std::atoi(c.at("sampling_resolution").c_str());
	if (has(c, "windowing")) {
		const std::string &w = c.at("windowing");
		if (has(c, "sigma"))
			sigma = (float)std::atof(c.at("sigma").c_str());
		if (w == "gauss")
			window = HCS_WINDOW_GAUSS;
		else if (w == "tukey")
			window = HCS_WINDOW_TUKEY;
		else if (w == "square")
			window = HCS_WINDOW_SQUARE; // unknown names fall back to none (flat_tactile_sensor.cpp:165-167)
	}
	int cfg_idx = owner_->configIndex(geomID);
	if (cfg_idx < 0)
		return false; // the sensor geom has no cs:: entry: it can never be part of a contact surface
	sensor_index_ = hcs_add_flat_sensor(owner_->context(), cfg_idx, resolution, sampling_resolution, window, sigma);
	if (sensor_index_ < 0) {
		std::fprintf(stderr, "[mujoco_contact_surface_sensors] %s\n", hcs_last_error(owner_->context()));
		return false;
	}
	hcs_sensor_dims(owner_->context(), sensor_index_, &cx, &cy);
	vGeoms = new mjvGeom[2 * cx * cy + 50];
	tactile_state_values_.assign((size_t)cx * cy, 0.0f);
	return true;
}

// flat_tactile_sensor.cpp:48-125.  As in the reference the taxel resolution is NOT changed by a request (:66 is commented
// out there); update_rate and visualize are host-side state, sampling_resolution / window / sigma go to the engine.
void FlatTactileSensor::dynamicParamCallback(DynamicFlatTactileConfig &config, uint32_t level, const mjModel *)
{
	if (level == (uint32_t)-1) { // when initializing fetch current config instead of overriding
		config.update_rate         = 1.0 / updatePeriod;
		config.visualize           = visualize;
		config.resolution          = resolution;
		config.sampling_resolution = sampling_resolution;
		config.window              = window;
		config.sigma               = sigma;
		return;
	}
	std::lock_guard<std::mutex> lock(pause_mutex); // the update loop holds it while it runs
	updateRate          = config.update_rate;
	updatePeriod        = 1.0 / config.update_rate;
	visualize           = config.visualize;
	sampling_resolution = config.sampling_resolution;
	sigma               = (float)config.sigma;
	window              = config.window >= 1 && config.window <= 3 ? config.window : HCS_WINDOW_NONE;
	if (hcs_update_flat_sensor(owner_->context(), sensor_index_, sampling_resolution, window, sigma) != HCS_OK)
		std::fprintf(stderr, "[mujoco_contact_surface_sensors] %s\n", hcs_last_error(owner_->context()));
	n_vGeom = 0;
	std::fill(tactile_state_values_.begin(), tactile_state_values_.end(), 0.0f); // channel.values.resize(cx * cy), :121-124
}

// flat_tactile_sensor.cpp:216-221 -> bvh_update (:262-402), computed on the GPU by the tactile kernels
void FlatTactileSensor::internal_update(const mjModel *m, mjData *d, const std::vector<GeomCollisionPtr> &)
{
	if (!owner_->finalized()) { // no contact pair seen yet: zero image (flat_tactile_sensor.cpp:290-296)
		std::fill(tactile_state_values_.begin(), tactile_state_values_.end(), 0.0f);
		return;
	}
	if (hcs_get_sensor_image(owner_->context(), sensor_index_, tactile_state_values_.data()) != HCS_OK)
		return; // this step ran without sensors (request arrived mid-step): keep the previous message
	if (visualize) { // render_tiles, flat_tactile_sensor.cpp:223-260
		const mjtNum *rot = d->geom_xmat + 9 * geomID, *xp = d->geom_xpos + 3 * geomID;
		const mjtNum *gs  = m->geom_size + 3 * geomID;
		float peak        = 1e-9f;
		for (float v : tactile_state_values_)
			peak = std::max(peak, std::fabs(v));
		for (int x = 0; x < cx; ++x)
			for (int y = 0; y < cy; ++y) {
				int idx = x + cy * y;
				if (idx >= cx * cy)
					continue;
				float ps            = std::fabs(tactile_state_values_[idx]) / peak;
				const float rgba[4] = { ps, 0, 1.f - ps, 0.8f };
				mjtNum l[3] = { -gs[0] + x * resolution + resolution / 2, -gs[1] + y * resolution + resolution / 2, gs[2] };
				mjtNum pos[3];
				for (int r = 0; r < 3; ++r)
					pos[r] = rot[3 * r] * l[0] + rot[3 * r + 1] * l[1] + rot[3 * r + 2] * l[2] + xp[r];
				mjtNum size[3] = { resolution / 2, resolution / 2, 0.0005 };
				mjv_initGeom(vGeoms + n_vGeom++, mjGEOM_BOX, size, pos, rot, rgba);
			}
	}
}

static std::vector<double> parse_numbers(const std::string &text)
{
	std::vector<double> out;
	std::string tok;
	auto flush = [&] {
		if (!tok.empty())
			out.push_back(std::atof(tok.c_str())), tok.clear();
	};
	for (char ch : text) {
		if (std::isdigit((unsigned char)ch) || ch == '-' || ch == '+' || ch == '.' || ch == 'e' || ch == 'E')
			tok.push_back(ch);
		else
			flush();
	}
	flush();
	return out;
}

// curved_sensor.cpp:111-380
bool CurvedSensor::load(const mjModel *m, mjData *d)
{
	const PluginConfig &c = rosparam_config_;
	if (!(TactileSensorBase::load(m, d) && has(c, "taxels") && has(c, "method") && has(c, "include_margin") &&
	      has(c, "sample_resolution")) ||
	    !owner_ || !owner_->context())
		return false;
	include_margin    = std::atof(c.at("include_margin").c_str());
	sample_resolution = std::atof(c.at("sample_resolution").c_str());
	const std::string &method = c.at("method"); // every method evaluates the same weighted ray sum (:443-479)
	if (method != "closest" && method != "weighted" && method != "mean" && method != "squared")
		return false;
	taxel_pos_ = parse_numbers(c.at("taxels"));
	if (taxel_pos_.empty() || taxel_pos_.size() % 3 != 0)
		return false;
	if (has(c, "normals")) {
		taxel_nrm_ = parse_numbers(c.at("normals"));
		if (taxel_nrm_.size() != taxel_pos_.size())
			return false;
	}
	if (m->geom_type[geomID] != mjGEOM_MESH || m->geom_dataid[geomID] < 0)
		return false; // the reference samples mesh geoms only (:241 "TODO implement methods to sample on primitive ...")
	int cfg_idx = owner_->configIndex(geomID);
	if (cfg_idx < 0)
		return false;
	// surface samples: area-weighted, fixed seed (see the class comment)
	const int did = m->geom_dataid[geomID];
	const int nf  = m->mesh_facenum[did];
	const float *mv = m->mesh_vert + 3 * m->mesh_vertadr[did];
	const int *mf   = m->mesh_face + 3 * m->mesh_faceadr[did];
	std::vector<double> cum(nf);
	double total = 0;
	auto vert = [&](int v, double out[3]) {
		for (int a = 0; a < 3; ++a)
			out[a] = mv[3 * v + a];
	};
	for (int f = 0; f < nf; ++f) {
		double a[3], b[3], cc[3];
		vert(mf[3 * f], a), vert(mf[3 * f + 1], b), vert(mf[3 * f + 2], cc);
		double u[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, v[3] = { cc[0] - a[0], cc[1] - a[1], cc[2] - a[2] };
		double n[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
		total += 0.5 * std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
		cum[f] = total;
	}
	// curved_sensor.cpp:276-283: sampleNum = (int)sample_resolution * total_area (the cast comes first: 0 for the shipped
	// sample_resolution 0.0005), then vcg::tri::PoissonSampling(mesh, points, sampleNum, radius = 0).  vcglib (un-vendored;
	// restated from its documented behaviour): a Monte-Carlo pool of max(10000, 40 * sampleNum) area-weighted uniform
	// points is pruned to a Poisson-disk set of radius sqrt(area / (0.7 pi sampleNum)); with sampleNum == 0 the radius
	// is 0 and the whole pool of 10000 points survives.  Which points the pruning keeps is vcglib's business (its
	// generator, seed 42, and its visiting order); here: splitmix64 seed 42, greedy in generation order.
	const int sample_num  = (int)((int)sample_resolution * total);
	const long n_samples  = std::min<long>(std::max<long>(10000, 40L * sample_num), 4000000);
	const double p_radius = sample_num > 0 ? std::sqrt(total / (0.7 * 3.14159265358979323846 * sample_num)) : 0.0;
	std::unordered_map<uint64_t, std::vector<int>> grid; // Poisson-disk pruning: accepted samples by cell of size radius
	auto cell_of = [&](const double *p, int o[3]) {
		for (int a = 0; a < 3; ++a)
			o[a] = (int)std::floor(p[a] / p_radius);
	};
	auto cell_key = [](int x, int y, int z) {
		return ((uint64_t)(uint32_t)(x + (1 << 20)) << 42) ^ ((uint64_t)(uint32_t)(y + (1 << 20)) << 21) ^ (uint64_t)(uint32_t)(z + (1 << 20));
	};
	uint64_t state = 42; // splitmix64
	auto uniform = [&]() {
		uint64_t z = (state += 0x9e3779b97f4a7c15ULL);
		z          = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
		z          = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
		z ^= z >> 31;
		return (double)(z >> 11) * (1.0 / 9007199254740992.0);
	};
	for (long k = 0; k < n_samples; ++k) {
		double r = uniform() * total;
		int f    = (int)(std::lower_bound(cum.begin(), cum.end(), r) - cum.begin());
		f        = std::min(f, nf - 1);
		double a[3], b[3], cc[3];
		vert(mf[3 * f], a), vert(mf[3 * f + 1], b), vert(mf[3 * f + 2], cc);
		double s0 = 1.0 - std::sqrt(uniform()), s1 = (1.0 - s0) * uniform();
		double u[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, v[3] = { cc[0] - a[0], cc[1] - a[1], cc[2] - a[2] };
		double n[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
		double l    = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
		double p[3];
		for (int ax = 0; ax < 3; ++ax)
			p[ax] = s0 * a[ax] + (1 - s0 - s1) * b[ax] + s1 * cc[ax];
		if (p_radius > 0) { // keep the point only if no accepted point lies within the disk radius
			int c0[3];
			cell_of(p, c0);
			bool free_spot = true;
			for (int dx = -1; dx <= 1 && free_spot; ++dx)
				for (int dy = -1; dy <= 1 && free_spot; ++dy)
					for (int dz = -1; dz <= 1 && free_spot; ++dz) {
						auto it = grid.find(cell_key(c0[0] + dx, c0[1] + dy, c0[2] + dz));
						if (it == grid.end())
							continue;
						for (int j : it->second) {
							const double *q = &sample_pos_[3 * (size_t)j];
							double d2 = (p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]);
							if (d2 < p_radius * p_radius) {
								free_spot = false;
								break;
							}
						}
					}
			if (!free_spot)
				continue;
			grid[cell_key(c0[0], c0[1], c0[2])].push_back((int)(sample_pos_.size() / 3));
		}
		for (int ax = 0; ax < 3; ++ax) {
			sample_pos_.push_back(p[ax]);
			sample_nrm_.push_back(l > 0 ? n[ax] / l : 0.0);
		}
	}
	const int n_taxels = (int)taxel_pos_.size() / 3;
	sensor_index_ = hcs_add_curved_sensor(owner_->context(), cfg_idx, n_taxels, taxel_pos_.data(),
	                                      taxel_nrm_.empty() ? nullptr : taxel_nrm_.data(), (int)(sample_pos_.size() / 3),
	                                      sample_pos_.data(), sample_nrm_.data(), include_margin);
	if (sensor_index_ < 0) {
		std::fprintf(stderr, "[mujoco_contact_surface_sensors] %s\n", hcs_last_error(owner_->context()));
		return false;
	}
	tactile_state_values_.assign((size_t)n_taxels, 0.0f);
	return true;
}

// curved_sensor.cpp:388-481, computed on the GPU by the curved-sensor kernels
void CurvedSensor::internal_update(const mjModel *, mjData *, const std::vector<GeomCollisionPtr> &)
{
	if (!owner_->finalized()) { // no contact pair seen yet: zeros (:446-450)
		std::fill(tactile_state_values_.begin(), tactile_state_values_.end(), 0.0f);
		return;
	}
	hcs_get_curved_values(owner_->context(), sensor_index_, tactile_state_values_.data());
}

// taxel_sensor.cpp:45-156
bool TaxelSensor::load(const mjModel *m, mjData *d)
{
	const PluginConfig &c = rosparam_config_;
	if (!(TactileSensorBase::load(m, d) && has(c, "taxels") && has(c, "method") && has(c, "include_margin") &&
	      has(c, "sample_resolution")) ||
	    !owner_ || !owner_->context())
		return false;
	include_margin    = std::atof(c.at("include_margin").c_str());
	sample_resolution = std::atof(c.at("sample_resolution").c_str());
	int sample_method = 0; // taxel_sensor.cpp:54-67
	if (has(c, "sample_method")) {
		const std::string &sm = c.at("sample_method");
		if (sm == "default")
			sample_method = 0;
		else if (sm == "area_importance")
			sample_method = 1;
		else {
			std::fprintf(stderr, "[mujoco_contact_surface_sensors] Could not find any match for sample_method: %s\n", sm.c_str());
			return false;
		}
	}
	const std::string &ms = c.at("method");
	int method = ms == "closest" ? 0 : ms == "weighted" ? 1 : ms == "mean" ? 2 : ms == "squared" ? 3 : -1;
	if (method < 0)
		return false;
	taxel_pos_ = parse_numbers(c.at("taxels"));
	if (taxel_pos_.empty() || taxel_pos_.size() % 3 != 0)
		return false;
	int cfg_idx = owner_->configIndex(geomID);
	if (cfg_idx < 0)
		return false;
	sensor_index_ = hcs_add_taxel_sensor(owner_->context(), cfg_idx, (int)taxel_pos_.size() / 3, taxel_pos_.data(),
	                                     include_margin, sample_resolution, method, visualize ? 1 : 0, sample_method);
	if (sensor_index_ < 0) {
		std::fprintf(stderr, "[mujoco_contact_surface_sensors] %s\n", hcs_last_error(owner_->context()));
		return false;
	}
	tactile_state_values_.assign(taxel_pos_.size() / 3, 0.0f);
	if (has(c, "visualize_max_pressure"))
		max_pressure = std::atof(c.at("visualize_max_pressure").c_str());
	if (visualize)
		vGeoms = new mjvGeom[taxel_pos_.size() / 3 + 1];
	return true;
}

// taxel_sensor.cpp:158-478, computed on the GPU by the taxel-sensor kernels
void TaxelSensor::internal_update(const mjModel *, mjData *d, const std::vector<GeomCollisionPtr> &)
{
	if (!owner_->finalized()) // no contact pair seen yet: zeros (:455-477)
		std::fill(tactile_state_values_.begin(), tactile_state_values_.end(), 0.0f);
	else
		hcs_get_taxel_values(owner_->context(), sensor_index_, tactile_state_values_.data());
	if (visualize && vGeoms) {
		// one sphere per taxel at its world position, colour and size by pressure / visualize_max_pressure
		// (:447-455 and the same block in every method; without samples: blue spheres of 0.5 mm, :470-476, which is what
		// scale 0 gives).  The 0.1 mm markers of the in-range sample points (:397-402) stay on the GPU and are not drawn.
		const int id = geomID;
		const mjtNum *R = d->geom_xmat + 9 * id, *xp = d->geom_xpos + 3 * id;
		const int n = (int)taxel_pos_.size() / 3;
		for (int i = 0; i < n; ++i) {
			const double *t = &taxel_pos_[3 * (size_t)i];
			mjtNum pos[3];
			for (int r = 0; r < 3; ++r)
				pos[r] = R[3 * r] * t[0] + R[3 * r + 1] * t[1] + R[3 * r + 2] * t[2] + xp[r];
			const double pressure = tactile_state_values_[i];
			const float scale     = (float)(std::min(std::max(pressure, 0.0), max_pressure) / max_pressure);
			const float color[4]  = { scale, 0, 1.0f - scale, 1 };
			const mjtNum sz       = 0.0005 + scale * 0.002;
			const mjtNum size[3]  = { sz, sz, sz };
			mjv_initGeom(vGeoms + n_vGeom++, mjGEOM_SPHERE, size, pos, nullptr, color);
		}
	}
}

} // namespace sensors
} // namespace mujoco_ros::contact_surfaces
