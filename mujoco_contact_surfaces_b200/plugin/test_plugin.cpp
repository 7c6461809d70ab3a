// Drives the host adapter the way mujoco_ros::MujocoEnv drives the reference plugin:
// load(), then per step: MuJoCo's collision pass calls mjCOLLISIONFUNC[t1][t2] for every geom pair, then the
// passive callback runs.  Worlds are shim re-creations of the reference's example worlds
//   mujoco_contact_surfaces/assets/sphere_on_box_world.xml          (config 1)
//   mujoco_contact_surface_sensors/assets/myrmex_box_world.xml + config/myrmex_sensor.yaml  (config 2)
// Prints one JSON object per scenario; tests/test_plugin_adapter.py checks it against the oracle.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "contact_surfaces_plugin.h"

using namespace mujoco_ros::contact_surfaces;

struct ShimWorld {
	mjModel m{};
	mjData d{};
	std::vector<int> geom_type, geom_bodyid, geom_dataid, numeric_adr, numeric_size, text_adr, text_size, body_dofadr;
	std::vector<double> geom_size, numeric_data, xpos, xmat, xipos, qfrc, vel6;
	std::vector<char> text_data;
	std::vector<std::string> gnames, nnames, tnames;
	std::vector<const char *> gptr, nptr, tptr;

	int add_body(bool free_joint, const double com[3])
	{
		int b = (int)body_dofadr.size();
		body_dofadr.push_back(free_joint ? m.nv : -1);
		if (free_joint)
			m.nv += 6;
		xipos.insert(xipos.end(), com, com + 3);
		return b;
	}
	std::vector<float> mesh_vert;
	std::vector<int> mesh_face, mesh_vertadr, mesh_vertnum, mesh_faceadr, mesh_facenum;
	int add_mesh(const std::vector<float> &v, const std::vector<int> &f)
	{
		mesh_vertadr.push_back((int)mesh_vert.size() / 3), mesh_vertnum.push_back((int)v.size() / 3);
		mesh_faceadr.push_back((int)mesh_face.size() / 3), mesh_facenum.push_back((int)f.size() / 3);
		mesh_vert.insert(mesh_vert.end(), v.begin(), v.end());
		mesh_face.insert(mesh_face.end(), f.begin(), f.end());
		return (int)mesh_vertadr.size() - 1;
	}
	int add_geom(const std::string &name, int type, int body, const double size[3], const double pos[3], const double mat[9],
	             int dataid = -1)
	{
		geom_type.push_back(type), geom_bodyid.push_back(body), geom_dataid.push_back(dataid);
		geom_size.insert(geom_size.end(), size, size + 3);
		xpos.insert(xpos.end(), pos, pos + 3);
		xmat.insert(xmat.end(), mat, mat + 9);
		vel6.insert(vel6.end(), 6, 0.0);
		gnames.push_back(name);
		return (int)geom_type.size() - 1;
	}
	void add_numeric(const std::string &name, std::vector<double> data)
	{
		nnames.push_back(name);
		numeric_adr.push_back((int)numeric_data.size());
		numeric_size.push_back((int)data.size());
		numeric_data.insert(numeric_data.end(), data.begin(), data.end());
	}
	void add_text(const std::string &name, const std::string &data)
	{
		tnames.push_back(name);
		text_adr.push_back((int)text_data.size());
		text_size.push_back((int)data.size());
		text_data.insert(text_data.end(), data.begin(), data.end());
	}
	void finish()
	{
		for (auto &s : gnames) gptr.push_back(s.c_str());
		for (auto &s : nnames) nptr.push_back(s.c_str());
		for (auto &s : tnames) tptr.push_back(s.c_str());
		m.ngeom = (int)geom_type.size(), m.nbody = (int)body_dofadr.size();
		m.nnumeric = (int)nnames.size(), m.ntext = (int)tnames.size();
		m.geom_type = geom_type.data(), m.geom_bodyid = geom_bodyid.data(), m.geom_dataid = geom_dataid.data();
		m.geom_size = geom_size.data();
		m.numeric_adr = numeric_adr.data(), m.numeric_size = numeric_size.data(), m.numeric_data = numeric_data.data();
		m.text_adr = text_adr.data(), m.text_size = text_size.data(), m.text_data = text_data.data();
		m.geom_names = gptr.data(), m.numeric_names = nptr.data(), m.text_names = tptr.data();
		m.body_dofadr = body_dofadr.data();
		m.mesh_vert = mesh_vert.data(), m.mesh_face = mesh_face.data();
		m.mesh_vertadr = mesh_vertadr.data(), m.mesh_vertnum = mesh_vertnum.data();
		m.mesh_faceadr = mesh_faceadr.data(), m.mesh_facenum = mesh_facenum.data();
		qfrc.assign(m.nv, 0.0);
		d.geom_xpos = xpos.data(), d.geom_xmat = xmat.data(), d.xipos = xipos.data();
		d.qfrc_passive = qfrc.data(), d.geom_vel6 = vel6.data();
		d.time = 0;
	}
	// mj_collision: every geom pair, ordered by geom type like mj_collideGeoms
	void collision_pass()
	{
		mjContact con[8];
		for (int a = 0; a < m.ngeom; ++a)
			for (int b = a + 1; b < m.ngeom; ++b) {
				if (geom_bodyid[a] == geom_bodyid[b])
					continue;
				int g1 = a, g2 = b;
				if (geom_type[g1] > geom_type[g2])
					std::swap(g1, g2);
				mjCOLLISIONFUNC[geom_type[g1]][geom_type[g2]](&m, &d, con, g1, g2, 0);
			}
	}
};

static const double I3[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };

static void rot_zyx(double yaw, double pitch, double roll, double R[9])
{
	double cz = cos(yaw), sz = sin(yaw), cy = cos(pitch), sy = sin(pitch), cx = cos(roll), sx = sin(roll);
	double Rz[9] = { cz, -sz, 0, sz, cz, 0, 0, 0, 1 }, Ry[9] = { cy, 0, sy, 0, 1, 0, -sy, 0, cy }, Rx[9] = { 1, 0, 0, 0, cx, -sx, 0, sx, cx };
	double T[9];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			T[3 * i + j] = 0;
			for (int k = 0; k < 3; ++k)
				T[3 * i + j] += Rz[3 * i + k] * Ry[3 * k + j];
		}
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			R[3 * i + j] = 0;
			for (int k = 0; k < 3; ++k)
				R[3 * i + j] += T[3 * i + k] * Rx[3 * k + j];
		}
}

static void print_vec(const char *key, const double *v, int n, bool last = false)
{
	std::printf("\"%s\": [", key);
	for (int i = 0; i < n; ++i)
		std::printf("%s%.17g", i ? ", " : "", v[i]);
	std::printf("]%s", last ? "" : ", ");
}

static int scenario_sphere_on_box()
{
	ShimWorld w;
	const double zero[3] = { 0, 0, 0 };
	double box_pos[3] = { 0, 0, 0.1 }, sph_pos[3] = { 0.012, -0.02, 0.2 + 0.08 - 0.012 };
	double R[9];
	rot_zyx(0.4, -0.3, 0.7, R);
	int b0 = w.add_body(false, zero), b1 = w.add_body(true, box_pos), b2 = w.add_body(true, sph_pos);
	double s_plane[3] = { 0, 0, 1 }, s_box[3] = { 0.1, 0.1, 0.1 }, s_sph[3] = { 0.08, 0, 0 };
	w.add_geom("ground", mjGEOM_PLANE, b0, s_plane, zero, I3);
	w.add_geom("box0", mjGEOM_BOX, b1, s_box, box_pos, I3);
	int gs = w.add_geom("sphere0", mjGEOM_SPHERE, b2, s_sph, sph_pos, R);
	w.add_text("cs::HydroelasticContactRepresentation", "kPolygon");
	w.add_numeric("cs::VisualizeSurfaces", { 1 });
	w.add_numeric("cs::box0", { 0, 1.0, 0.1, 0.3, 0.3 });
	w.add_numeric("cs::sphere0", { 5e4, 5.0, 0.05, 0.3, 0.3 });
	w.finish();
	double v[6] = { 0.3, -0.2, 0.1, 0.02, 0.01, -0.05 };
	std::memcpy(&w.vel6[6 * gs], v, sizeof v);

	MujocoContactSurfacesPlugin plugin;
	if (!plugin.load(&w.m, &w.d)) {
		std::printf("{\"scenario\": \"sphere_on_box\", \"error\": \"load failed\"}\n");
		return 1;
	}
	int n_geoms_seen = 0;
	for (int step = 0; step < 3; ++step) { // identical steps: results must not depend on history
		std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
		w.collision_pass();
		plugin.passiveCallback(&w.m, &w.d);
		w.d.time += 0.001;
	}
	static mjvGeom scene_geoms[4096];
	mjvScene scene{ 4096, 0, scene_geoms };
	plugin.renderCallback(&w.m, &w.d, &scene);
	n_geoms_seen = scene.ngeom;
	double outline = 0; // the contact surface is drawn as one connector per face edge (plugin.cpp:525-555)
	int n_cyl      = 0;
	for (int i = 0; i < scene.ngeom; ++i)
		if (scene.geoms[i].type == mjGEOM_CYLINDER) {
			outline += 2.0 * scene.geoms[i].size[2];
			++n_cyl;
		}
	std::printf("{\"scenario\": \"sphere_on_box\", ");
	print_vec("box_pos", box_pos, 3);
	print_vec("sphere_pos", sph_pos, 3);
	print_vec("sphere_mat", R, 9);
	print_vec("sphere_vel6", v, 6);
	std::printf("\"vgeoms\": %d, \"connectors\": %d, \"outline_length\": %.9g, ", n_geoms_seen, n_cyl, outline);
	print_vec("qfrc_passive", w.qfrc.data(), (int)w.qfrc.size(), true);
	std::printf("}\n");
	return 0;
}

static int scenario_myrmex()
{
	ShimWorld w;
	const double zero[3] = { 0, 0, 0 };
	double foam_pos[3] = { 0, 0, 0.033 };
	double R[9];
	rot_zyx(0.6, 0.01, -0.015, R);
	// lowest corner of the tilted 0.2 m cube 3 mm below the foam top (z = 0.053)
	double low = 0;
	for (int sx = -1; sx <= 1; sx += 2)
		for (int sy = -1; sy <= 1; sy += 2)
			for (int sz = -1; sz <= 1; sz += 2)
				low = std::fmin(low, 0.1 * (R[6] * sx + R[7] * sy + R[8] * sz));
	double box_pos[3] = { 0.02, -0.01, 0.053 - 0.003 - low };
	int b0 = w.add_body(false, zero), b1 = w.add_body(true, box_pos);
	double s_plane[3] = { 0, 0, 1 }, s_box[3] = { 0.1, 0.1, 0.1 }, s_foam[3] = { 0.2, 0.2, 0.02 };
	w.add_geom("ground", mjGEOM_PLANE, b0, s_plane, zero, I3);
	w.add_geom("box1", mjGEOM_BOX, b1, s_box, box_pos, R);
	w.add_geom("myrmex_foam", mjGEOM_BOX, b0, s_foam, foam_pos, I3);
	w.add_text("cs::HydroelasticContactRepresentation", "kTriangle");
	w.add_numeric("cs::VisualizeSurfaces", { 0 });
	w.add_numeric("cs::ApplyContactSurfaceForces", { 1 });
	w.add_numeric("cs::box1", { 0, 1.0, 0.05, 0.3, 0.3 });
	w.add_numeric("cs::plate", { 0, 1.0, 0, 0.3, 0.3 }); // names a geom that is not in this world: ignored
	w.add_numeric("cs::myrmex_foam", { 5e4, 5.0, 0, 0.3, 0.3 });
	w.finish();

	MujocoContactSurfacesPlugin plugin;
	auto sensor = std::make_shared<sensors::FlatTactileSensor>();
	// config/myrmex_sensor.yaml:4
	PluginConfig cfg = { { "type", "mujoco_contact_surface_sensors/FlatTactileSensor" }, { "sensorName", "myrmex_sensor0" },
		                 { "geomName", "myrmex_foam" }, { "topicName", "/tactile_module_16x16_v2" }, { "updateRate", "50.0" },
		                 { "visualize", "True" }, { "use_parallel", "True" }, { "resolution", "0.025" },
		                 { "sampling_resolution", "20" } };
	plugin.addSurfacePlugin(sensor, cfg);
	if (!plugin.load(&w.m, &w.d)) {
		std::printf("{\"scenario\": \"myrmex_box\", \"error\": \"load failed\"}\n");
		return 1;
	}
	// 45 steps of 1 ms at 50 Hz: the sensor publishes at t = 0, 0.020, 0.040
	for (int step = 0; step < 45; ++step) {
		std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
		w.collision_pass();
		plugin.passiveCallback(&w.m, &w.d);
		w.d.time += 0.001;
	}
	// dynamic_reconfigure request (DynamicFlatTactile.cfg): 8 x 8 rays per taxel, gauss window, then one more update
	std::vector<double> img_before(sensor->lastMessage().begin(), sensor->lastMessage().end());
	sensors::DynamicFlatTactileConfig dyn;
	sensor->dynamicParamCallback(dyn, (uint32_t)-1, &w.m); // fetch
	const int fetched_sampling = dyn.sampling_resolution;
	dyn.sampling_resolution = 8, dyn.window = 1, dyn.sigma = 0.1;
	sensor->dynamicParamCallback(dyn, 0, &w.m);
	const int publishes_before = sensor->publishCount();
	for (int step = 0; step < 25 && sensor->publishCount() == publishes_before; ++step) {
		std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
		w.collision_pass();
		plugin.passiveCallback(&w.m, &w.d);
		w.d.time += 0.001;
	}
	std::vector<double> img_reconf(sensor->lastMessage().begin(), sensor->lastMessage().end());
	std::printf("{\"scenario\": \"myrmex_box\", \"fetched_sampling\": %d, ", fetched_sampling);
	print_vec("image_reconfigured", img_reconf.data(), (int)img_reconf.size());
	print_vec("box_pos", box_pos, 3);
	print_vec("box_mat", R, 9);
	std::printf("\"cx\": %d, \"cy\": %d, \"publishes\": %d, ", sensor->cx, sensor->cy, sensor->publishCount());
	print_vec("image", img_before.data(), (int)img_before.size());
	print_vec("qfrc_passive", w.qfrc.data(), (int)w.qfrc.size(), true);
	std::printf("}\n");
	return 0;
}

// CurvedSensor on a soft convex mesh geom (SENS/assets/fingertip_mocap.xml arrangement with a small octahedral tip)
static int scenario_curved_tip()
{
	ShimWorld w;
	const double zero[3] = { 0, 0, 0 };
	// octahedron with semi-axes (8, 6, 10) mm, outward winding
	std::vector<float> mv = { 0.008f, 0, 0, -0.008f, 0, 0, 0, 0.006f, 0, 0, -0.006f, 0, 0, 0, 0.010f, 0, 0, -0.010f };
	std::vector<int> mf   = { 0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5 };
	double box_pos[3] = { 0, 0, 0.025 };
	double R[9];
	rot_zyx(0.3, 0.25, -0.2, R);
	double low = 0; // lowest vertex of the rotated tip, 1.5 mm into the box top (z = 0.05)
	for (size_t v = 0; v < mv.size() / 3; ++v)
		low = std::fmin(low, R[6] * mv[3 * v] + R[7] * mv[3 * v + 1] + R[8] * mv[3 * v + 2]);
	double tip_pos[3] = { 0.004, -0.003, 0.05 - 0.0015 - low };
	int b0 = w.add_body(false, zero), b1 = w.add_body(true, tip_pos);
	double s_box[3] = { 0.025, 0.025, 0.025 };
	int did = w.add_mesh(mv, mf);
	w.add_geom("box_geom", mjGEOM_BOX, b0, s_box, box_pos, I3);
	w.add_geom("fingertip_geom", mjGEOM_MESH, b1, zero, tip_pos, R, did);
	w.add_text("cs::HydroelasticContactRepresentation", "kTriangle");
	w.add_numeric("cs::box_geom", { 0, 1.0, 0.01, 0.0, 0.0 });
	w.add_numeric("cs::fingertip_geom", { 5e4, 5.0, 0.0, 0.0, 0.0 });
	w.finish();

	MujocoContactSurfacesPlugin plugin;
	auto sensor = std::make_shared<sensors::CurvedSensor>();
	// keys of config/curved_fingertip.yaml; taxels on the lower faces of the tip
	PluginConfig cfg = { { "type", "mujoco_contact_surface_sensors/CurvedSensor" }, { "sensorName", "myrmex_fingertip" },
		                 { "geomName", "fingertip_geom" }, { "topicName", "/myrmex_fingertip" }, { "updateRate", "4.0" },
		                 { "include_margin", "0.006" }, { "method", "squared" }, { "sample_method", "area_importance" },
		                 { "sample_resolution", "0.0008" },
		                 { "taxels", "[[0.002, 0.0015, -0.005], [-0.002, 0.0015, -0.005], [0.002, -0.0015, -0.005], "
		                             "[-0.002, -0.0015, -0.005], [0, 0, 0.010]]" },
		                 { "normals", "[[0.5, 0.6, -0.6], [-0.5, 0.6, -0.6], [0.5, -0.6, -0.6], [-0.5, -0.6, -0.6], [0, 0, 1]]" } };
	plugin.addSurfacePlugin(sensor, cfg);
	if (!plugin.load(&w.m, &w.d)) {
		std::printf("{\"scenario\": \"curved_tip\", \"error\": \"load failed\"}\n");
		return 1;
	}
	for (int step = 0; step < 3; ++step) {
		std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
		w.collision_pass();
		plugin.passiveCallback(&w.m, &w.d);
		w.d.time += 0.001;
	}
	std::printf("{\"scenario\": \"curved_tip\", ");
	print_vec("tip_pos", tip_pos, 3);
	print_vec("tip_mat", R, 9);
	std::vector<double> mvd(mv.begin(), mv.end()), mfd(mf.begin(), mf.end());
	print_vec("mesh_vert", mvd.data(), (int)mvd.size());
	print_vec("mesh_face", mfd.data(), (int)mfd.size());
	print_vec("taxels", sensor->taxelPoints().data(), (int)sensor->taxelPoints().size());
	print_vec("normals", sensor->taxelNormals().data(), (int)sensor->taxelNormals().size());
	print_vec("sample_pos", sensor->samplePoints().data(), (int)sensor->samplePoints().size());
	print_vec("sample_nrm", sensor->sampleNormals().data(), (int)sensor->sampleNormals().size());
	std::printf("\"publishes\": %d, ", sensor->publishCount());
	std::vector<double> vals(sensor->lastMessage().begin(), sensor->lastMessage().end());
	print_vec("values", vals.data(), (int)vals.size(), true);
	std::printf("}\n");
	return 0;
}

// TaxelSensor with the keys of SENS/config/fingertip.yaml (method squared, sample_method area_importance, sample_resolution
// 0.001, include_margin 0.006) on the same small soft tip
static int scenario_taxel_tip()
{
	ShimWorld w;
	const double zero[3] = { 0, 0, 0 };
	std::vector<float> mv = { 0.008f, 0, 0, -0.008f, 0, 0, 0, 0.006f, 0, 0, -0.006f, 0, 0, 0, 0.010f, 0, 0, -0.010f };
	std::vector<int> mf   = { 0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5 };
	double box_pos[3] = { 0, 0, 0.025 };
	double R[9];
	rot_zyx(-0.4, 0.2, 0.15, R);
	double low = 0;
	for (size_t v = 0; v < mv.size() / 3; ++v)
		low = std::fmin(low, R[6] * mv[3 * v] + R[7] * mv[3 * v + 1] + R[8] * mv[3 * v + 2]);
	double tip_pos[3] = { -0.002, 0.005, 0.05 - 0.002 - low };
	int b0 = w.add_body(false, zero), b1 = w.add_body(true, tip_pos);
	double s_box[3] = { 0.025, 0.025, 0.025 };
	int did = w.add_mesh(mv, mf);
	w.add_geom("box_geom", mjGEOM_BOX, b0, s_box, box_pos, I3);
	w.add_geom("fingertip_geom", mjGEOM_MESH, b1, zero, tip_pos, R, did);
	w.add_text("cs::HydroelasticContactRepresentation", "kTriangle");
	w.add_numeric("cs::box_geom", { 0, 1.0, 0.01, 0.0, 0.0 });
	w.add_numeric("cs::fingertip_geom", { 5e4, 5.0, 0.0, 0.0, 0.0 });
	w.finish();

	MujocoContactSurfacesPlugin plugin;
	auto sensor = std::make_shared<sensors::TaxelSensor>();
	PluginConfig cfg = { { "type", "mujoco_contact_surface_sensors/TaxelSensor" }, { "sensorName", "myrmex_fingertip" },
		                 { "geomName", "fingertip_geom" }, { "topicName", "/myrmex_fingertip" }, { "updateRate", "4.0" },
		                 { "include_margin", "0.006" }, { "method", "squared" }, { "sample_method", "area_importance" },
		                 { "sample_resolution", "0.001" }, { "visualize", "True" }, { "visualize_max_pressure", "0.04" },
		                 { "taxels", "[[0.002, 0.0015, -0.005], [-0.002, 0.0015, -0.005], [0.002, -0.0015, -0.005], "
		                             "[-0.002, -0.0015, -0.005], [0, 0, 0.010]]" } };
	plugin.addSurfacePlugin(sensor, cfg);
	if (!plugin.load(&w.m, &w.d)) {
		std::printf("{\"scenario\": \"taxel_tip\", \"error\": \"load failed\"}\n");
		return 1;
	}
	for (int step = 0; step < 3; ++step) {
		std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
		w.collision_pass();
		plugin.passiveCallback(&w.m, &w.d);
		w.d.time += 0.001;
	}
	static mjvGeom scene_geoms[64];
	mjvScene scene{ 64, 0, scene_geoms };
	plugin.renderCallback(&w.m, &w.d, &scene);
	int n_spheres = 0;
	double max_size = 0;
	for (int i = 0; i < scene.ngeom; ++i)
		if (scene.geoms[i].type == mjGEOM_SPHERE) {
			++n_spheres;
			max_size = std::fmax(max_size, scene.geoms[i].size[0]);
		}
	std::printf("{\"scenario\": \"taxel_tip\", \"taxel_markers\": %d, \"max_marker_size\": %.9g, ", n_spheres, max_size);
	print_vec("tip_pos", tip_pos, 3);
	print_vec("tip_mat", R, 9);
	std::vector<double> mvd(mv.begin(), mv.end()), mfd(mf.begin(), mf.end());
	print_vec("mesh_vert", mvd.data(), (int)mvd.size());
	print_vec("mesh_face", mfd.data(), (int)mfd.size());
	std::printf("\"publishes\": %d, ", sensor->publishCount());
	std::vector<double> vals(sensor->lastMessage().begin(), sensor->lastMessage().end());
	print_vec("values", vals.data(), (int)vals.size(), true);
	std::printf("}\n");
	return 0;
}

// CurvedSensor::load with an integer-valued sample_resolution: (int)sample_resolution * area > 0, so the surface samples
// are a Poisson-disk set (curved_sensor.cpp:276-283)
static int scenario_curved_poisson()
{
	ShimWorld w;
	const double zero[3] = { 0, 0, 0 };
	std::vector<float> mv = { 0.008f, 0, 0, -0.008f, 0, 0, 0, 0.006f, 0, 0, -0.006f, 0, 0, 0, 0.010f, 0, 0, -0.010f };
	std::vector<int> mf   = { 0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5 };
	double box_pos[3] = { 0, 0, 0.025 }, tip_pos[3] = { 0, 0, 0.058 };
	int b0 = w.add_body(false, zero), b1 = w.add_body(true, tip_pos);
	double s_box[3] = { 0.025, 0.025, 0.025 };
	int did = w.add_mesh(mv, mf);
	w.add_geom("box_geom", mjGEOM_BOX, b0, s_box, box_pos, I3);
	w.add_geom("fingertip_geom", mjGEOM_MESH, b1, zero, tip_pos, I3, did);
	w.add_text("cs::HydroelasticContactRepresentation", "kTriangle");
	w.add_numeric("cs::box_geom", { 0, 1.0, 0.01, 0.0, 0.0 });
	w.add_numeric("cs::fingertip_geom", { 5e4, 5.0, 0.0, 0.0, 0.0 });
	w.finish();
	MujocoContactSurfacesPlugin plugin;
	auto sensor = std::make_shared<sensors::CurvedSensor>();
	PluginConfig cfg = { { "type", "mujoco_contact_surface_sensors/CurvedSensor" }, { "sensorName", "tip" },
		                 { "geomName", "fingertip_geom" }, { "topicName", "/tip" }, { "updateRate", "4.0" },
		                 { "include_margin", "0.006" }, { "method", "squared" }, { "sample_resolution", "3000000" },
		                 { "taxels", "[[0.002, 0.0015, -0.005], [0, 0, 0.010]]" } };
	plugin.addSurfacePlugin(sensor, cfg);
	if (!plugin.load(&w.m, &w.d)) {
		std::printf("{\"scenario\": \"curved_poisson\", \"error\": \"load failed\"}\n");
		return 1;
	}
	const std::vector<double> &p = sensor->samplePoints();
	const int n = (int)p.size() / 3;
	double dmin = 1e300;
	for (int i = 0; i < n; ++i)
		for (int j = i + 1; j < n; ++j) {
			double d2 = 0;
			for (int a = 0; a < 3; ++a)
				d2 += (p[3 * i + a] - p[3 * j + a]) * (p[3 * i + a] - p[3 * j + a]);
			dmin = std::fmin(dmin, d2);
		}
	double area = 0;
	for (size_t f = 0; f < mf.size() / 3; ++f) {
		double a[3], u[3], v[3];
		for (int k = 0; k < 3; ++k) {
			a[k] = mv[3 * mf[3 * f] + k];
			u[k] = mv[3 * mf[3 * f + 1] + k] - a[k];
			v[k] = mv[3 * mf[3 * f + 2] + k] - a[k];
		}
		double c[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
		area += 0.5 * std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
	}
	std::printf("{\"scenario\": \"curved_poisson\", \"n_samples\": %d, \"min_distance\": %.9g, \"area\": %.9g}\n", n,
	            std::sqrt(dmin), area);
	return 0;
}

// passiveCallback latency of the reference's own execution model: ONE mjData, one call per mj_step, no consumers of the
// per-face views (no sub-plugins, no visualisation).  Two worlds: config 1 (soft sphere on rigid box) and a config-4-like
// world (four soft objects on a rigid plane).  The pose changes a little every step; reported per step: the collision
// pass (dispatch only) and passiveCallback (poses in -> hcs_step -> wrenches applied).
static double now_us()
{
	return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int scenario_timing()
{
	const double zero[3] = { 0, 0, 0 };
	for (int world = 0; world < 2; ++world) {
		ShimWorld w;
		std::vector<int> movers;
		std::vector<double> base_z;
		int b0 = w.add_body(false, zero);
		if (world == 0) {
			double box_pos[3] = { 0, 0, 0.1 }, sph_pos[3] = { 0.012, -0.02, 0.2 + 0.08 - 0.012 };
			int b1 = w.add_body(true, box_pos), b2 = w.add_body(true, sph_pos);
			double s_box[3] = { 0.1, 0.1, 0.1 }, s_sph[3] = { 0.08, 0, 0 };
			w.add_geom("box0", mjGEOM_BOX, b1, s_box, box_pos, I3);
			movers.push_back(w.add_geom("sphere0", mjGEOM_SPHERE, b2, s_sph, sph_pos, I3));
			base_z.push_back(sph_pos[2]);
			w.add_numeric("cs::box0", { 0, 1.0, 0.1, 0.3, 0.3 });
			w.add_numeric("cs::sphere0", { 5e4, 5.0, 0.05, 0.3, 0.3 });
		} else {
			double s_plane[3] = { 0, 0, 1 };
			w.add_geom("ground", mjGEOM_PLANE, b0, s_plane, zero, I3);
			w.add_numeric("cs::ground", { 0, 1.0, 0, 0.5, 0.5 });
			const int types[4]    = { mjGEOM_SPHERE, mjGEOM_BOX, mjGEOM_ELLIPSOID, mjGEOM_CYLINDER };
			const double sizes[4][3] = { { 0.05, 0, 0 }, { 0.04, 0.05, 0.03 }, { 0.05, 0.03, 0.04 }, { 0.04, 0.05, 0 } };
			const double hz[4]       = { 0.05, 0.03, 0.04, 0.05 };
			const char *names[4]     = { "obj_sphere", "obj_box", "obj_ellipsoid", "obj_cylinder" };
			for (int k = 0; k < 4; ++k) {
				double pos[3] = { 0.3 * k, 0, hz[k] - 0.004 };
				int b         = w.add_body(true, pos);
				movers.push_back(w.add_geom(names[k], types[k], b, sizes[k], pos, I3));
				base_z.push_back(pos[2]);
				w.add_numeric(std::string("cs::") + names[k], { 5e4, 3.0, 0.02, 0.3, 0.3 });
			}
		}
		w.add_text("cs::HydroelasticContactRepresentation", "kPolygon");
		w.finish();
		MujocoContactSurfacesPlugin plugin;
		if (!plugin.load(&w.m, &w.d)) {
			std::printf("{\"scenario\": \"timing_%d\", \"error\": \"load failed\"}\n", world);
			return 1;
		}
		const int warm = 50, steps = 2000;
		double t_coll = 0, t_passive = 0, checksum = 0;
		for (int step = 0; step < warm + steps; ++step) {
			for (size_t k = 0; k < movers.size(); ++k)
				w.xpos[3 * movers[k] + 2] = base_z[k] - 0.002 * std::sin(0.01 * step + (double)k);
			std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
			double t0 = now_us();
			w.collision_pass();
			double t1 = now_us();
			plugin.passiveCallback(&w.m, &w.d);
			double t2 = now_us();
			if (step >= warm)
				t_coll += t1 - t0, t_passive += t2 - t1;
			for (double q : w.qfrc)
				checksum += std::fabs(q);
			w.d.time += 0.001;
		}
		std::printf("{\"scenario\": \"timing_%s\", \"steps\": %d, \"collision_pass_us\": %.3f, \"passive_callback_us\": %.3f, "
		            "\"qfrc_checksum\": %.9g}\n",
		            world == 0 ? "sphere_on_box" : "objects_on_plane", steps, t_coll / steps, t_passive / steps, checksum);
	}
	return 0;
}

// BatchedContactSurfaces: six mjData of the sphere-on-box world (different poses and velocities) through ONE multi-device
// context (two blocks), against six runs of the one-mjData plugin on the same poses: generalised forces bit for bit.
static int scenario_batched()
{
	const double zero[3] = { 0, 0, 0 };
	const int N = 6;
	std::vector<ShimWorld> worlds(N);
	std::vector<std::vector<double>> single(N);
	int gs = -1, gb = -1;
	for (int e = 0; e < N; ++e) {
		ShimWorld &w = worlds[e];
		double box_pos[3] = { 0, 0, 0.1 }, sph_pos[3] = { 0.01 * e - 0.02, 0.015 - 0.01 * e, 0.2 + 0.08 - 0.002 * (e + 1) };
		double R[9];
		rot_zyx(0.3 * e, -0.2 + 0.1 * e, 0.5 - 0.15 * e, R);
		w.add_body(false, zero);
		int b1 = w.add_body(true, box_pos), b2 = w.add_body(true, sph_pos);
		double s_box[3] = { 0.1, 0.1, 0.1 }, s_sph[3] = { 0.08, 0, 0 };
		gb = w.add_geom("box0", mjGEOM_BOX, b1, s_box, box_pos, I3);
		gs = w.add_geom("sphere0", mjGEOM_SPHERE, b2, s_sph, sph_pos, R);
		w.add_text("cs::HydroelasticContactRepresentation", "kPolygon");
		w.add_numeric("cs::box0", { 0, 1.0, 0.1, 0.3, 0.3 });
		w.add_numeric("cs::sphere0", { 5e4, 5.0, 0.05, 0.3, 0.3 });
		w.finish();
		double v[6] = { 0.3 - 0.1 * e, -0.2, 0.1 * e, 0.02, 0.01 * e, -0.05 };
		std::memcpy(&w.vel6[6 * gs], v, sizeof v);
	}
	for (int e = 0; e < N; ++e) { // the reference's way: one plugin instance per mjData
		ShimWorld &w = worlds[e];
		MujocoContactSurfacesPlugin plugin;
		if (!plugin.load(&w.m, &w.d)) {
			std::printf("{\"scenario\": \"batched\", \"error\": \"load failed\"}\n");
			return 1;
		}
		std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
		w.collision_pass();
		plugin.passiveCallback(&w.m, &w.d);
		single[e] = w.qfrc;
	}
	BatchedContactSurfaces batch;
	if (!batch.load(&worlds[0].m, N, { 0, 0 })) {
		std::printf("{\"scenario\": \"batched\", \"error\": \"batch load failed\"}\n");
		return 1;
	}
	batch.recordPair(gs, gb); // mj_collideGeoms order: sphere (2) before box (6)
	std::vector<mjData *> data;
	for (ShimWorld &w : worlds) {
		std::fill(w.qfrc.begin(), w.qfrc.end(), 0.0);
		data.push_back(&w.d);
	}
	batch.passiveCallback(&worlds[0].m, data.data(), N);
	double max_diff = 0, checksum = 0;
	for (int e = 0; e < N; ++e)
		for (size_t k = 0; k < single[e].size(); ++k) {
			max_diff = std::max(max_diff, std::fabs(single[e][k] - worlds[e].qfrc[k]));
			checksum += std::fabs(worlds[e].qfrc[k]);
		}
	std::printf("{\"scenario\": \"batched\", \"n\": %d, \"blocks\": %d, \"pairs\": %d, \"max_abs_diff\": %.17g, \"checksum\": %.9g}\n", N,
	            hcs_multi_n_blocks(batch.context()), batch.numPairs(), max_diff, checksum);
	return 0;
}

int main()
{
	int rc = scenario_sphere_on_box();
	rc |= scenario_myrmex();
	rc |= scenario_curved_tip();
	rc |= scenario_taxel_tip();
	rc |= scenario_curved_poisson();
	rc |= scenario_batched();
	rc |= scenario_timing();
	return rc;
}
