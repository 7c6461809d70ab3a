// Host-side drop-in for the reference's plugin classes, backed by libhcs_b200 (include/hcs.h).
//
// Same names, virtuals, argument meaning and error behaviour as
//   mujoco_contact_surfaces/include/mujoco_contact_surfaces/common_types.h:48-76      (PointCollision, GeomCollision)
//   mujoco_contact_surfaces/include/mujoco_contact_surfaces/plugin_utils.h:47-128     (SurfacePlugin)
//   mujoco_contact_surfaces/include/mujoco_contact_surfaces/mujoco_contact_surfaces_plugin.h:141-266
//   mujoco_contact_surface_sensors/include/mujoco_contact_surface_sensors/{tactile_sensor_base,flat_tactile_sensor}.h
// minus ROS (absent here): rosparam/XmlRpc configuration becomes PluginConfig (string -> string), ROS time
// becomes mjData.time, the TactileState message becomes a float vector with the same index order.
// The base class mujoco_ros::MujocoPlugin is replaced by the stub below with the same virtual surface.
#pragma once
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/hcs.h"
#include "mj_shim.h"

namespace mujoco_ros {

// virtual surface of mujoco_ros::MujocoPlugin used by the reference
class MujocoPlugin
{
public:
	virtual ~MujocoPlugin() {}
	virtual bool load(const mjModel *m, mjData *d) = 0;
	virtual void reset()                             = 0;
	virtual void controlCallback(const mjModel *, mjData *) {}
	virtual void passiveCallback(const mjModel *, mjData *) {}
	virtual void renderCallback(const mjModel *, mjData *, mjvScene *) {}
	virtual void lastStageCallback(const mjModel *, mjData *) {}
	virtual void onGeomChanged(const mjModel *, mjData *, const int) {}
};

namespace contact_surfaces {

const int MAX_VGEOM = 10000; // common_types.h:45

typedef std::map<std::string, std::string> PluginConfig; // stands in for XmlRpc::XmlRpcValue structs

// What a sub-plugin may read of a contact surface (the reference hands out drake::geometry::ContactSurface).
struct ContactSurfaceView {
	bool is_triangle = false;
	std::vector<double> triangles; // kTriangle: 12 doubles per face (9 world coordinates + 3 vertex pressures)
	double total_area = 0;
	double centroid[3] = { 0, 0, 0 };
	int num_faces = 0;
	// per PointCollision (index = PointCollision::face): the normal force fn of plugin.cpp:470 and, when the surfaces
	// are drawn (cs::VisualizeSurfaces), the face's world vertices, HCS_FACE_VERTEX_STRIDE doubles each
	std::vector<double> face_fn;
	std::vector<int> face_nverts;
	std::vector<double> face_vertices;
};

// common_types.h:48-56
typedef struct PointCollision
{
	double p[3];
	double n[3];
	double fn0;
	double stiffness;
	double damping;
	int face;
} PointCollision;

// common_types.h:59-66
typedef struct GeomCollision
{
	std::vector<PointCollision> pointCollisions;
	std::shared_ptr<ContactSurfaceView> s;
	int g1;
	int g2;
	GeomCollision(int g1, int g2, ContactSurfaceView *s) : s(s), g1(g1), g2(g2) {}
} GeomCollision;

typedef std::shared_ptr<GeomCollision> GeomCollisionPtr;
class SurfacePlugin;
typedef std::shared_ptr<SurfacePlugin> SurfacePluginPtr;
class MujocoContactSurfacesPlugin;

// plugin_utils.h:47-128
class SurfacePlugin
{
public:
	virtual ~SurfacePlugin() {}
	void init(const PluginConfig &config, const std::string &nh_namespace)
	{
		rosparam_config_ = config;
		namespace_       = nh_namespace;
	}
	bool safe_load(const mjModel *m, mjData *d)
	{
		loading_successful_ = load(m, d);
		return loading_successful_;
	}
	void safe_reset()
	{
		if (loading_successful_)
			reset();
	}
	virtual void update(const mjModel *, mjData *, const std::vector<GeomCollisionPtr> &) {}
	virtual void renderCallback(const mjModel *, mjData *, mjvScene *) {}
	// B200 extension: the owning plugin (and through it the GPU context) is known before load()
	void attach(MujocoContactSurfacesPlugin *owner) { owner_ = owner; }
	// B200 extension: does the plugin need sensor output in the step that is about to run?
	virtual bool wantsSensorUpdate(const mjData *) { return false; }

protected:
	virtual bool load(const mjModel *m, mjData *d) = 0;
	virtual void reset()                             = 0;
	SurfacePlugin() {}
	PluginConfig rosparam_config_;
	std::string namespace_;
	MujocoContactSurfacesPlugin *owner_ = nullptr;

private:
	bool loading_successful_ = false;
};

typedef enum _contactType { RIGID, SOFT } contactType;

// mujoco_contact_surfaces_plugin.h:141-213 without the Drake objects (they live on the GPU now)
struct ContactProperties
{
	int mujoco_geom_id;
	int drake_id; // configuration index = hcs geom index
	std::string geom_name;
	contactType contact_type;
	double hydroelastic_modulus, dissipation, static_friction, dynamic_friction, resolution_hint;
};

const std::string PREFIX = "cs::";

class MujocoContactSurfacesPlugin : public mujoco_ros::MujocoPlugin
{
public:
	MujocoContactSurfacesPlugin();
	virtual ~MujocoContactSurfacesPlugin();

	bool load(const mjModel *m, mjData *d) override;
	void reset() override;
	void passiveCallback(const mjModel *model, mjData *data) override;
	void renderCallback(const mjModel *model, mjData *data, mjvScene *scene) override;
	void onGeomChanged(const mjModel *model, mjData *data, const int geom_id) override;

	int collision_cb(const mjModel *m, const mjData *d, mjContact *con, int g1, int g2, mjtNum margin);
	std::vector<SurfacePluginPtr> getPlugins() { return plugins; }

	// configuration that the reference reads from rosparam "SurfacePlugins" (plugin_utils.cpp:41-119)
	void addSurfacePlugin(SurfacePluginPtr plugin, const PluginConfig &config);
	// engine access for GPU-resident sub-plugins
	hcs_ctx *context() { return ctx_; }
	int configIndex(int mujoco_geom_id) const;
	bool finalized() const { return finalized_; }
	int device = 0;

protected:
	const mjModel *m_ = nullptr;
	mjData *d_        = nullptr;
	bool visualizeContactSurfaces  = false;
	bool applyContactSurfaceForces = true;
	std::vector<GeomCollisionPtr> geomCollisions;

private:
	mjvGeom *vGeoms = new mjvGeom[MAX_VGEOM];
	int n_vGeom     = 0;
	double running_scale = 3., current_scale = 0.;
	int hydroelastic_contact_representation = HCS_REP_TRIANGLE; // reference default (plugin.h:251)
	std::map<int, std::shared_ptr<ContactProperties>> contactProperties;
	std::vector<int> cfg_to_mj; // configuration index -> mujoco geom id

	void parseMujocoCustomFields(const mjModel *m);
	void initCollisionFunction();
	void ensurePairs();
	void evaluateAndApply(const mjModel *m, mjData *d, bool with_sensors);
	void buildGeomCollisions();
	void visualizeMeshElement(int face, const ContactSurfaceView &mesh, double fn); // plugin.cpp:525-555

	hcs_ctx *ctx_   = nullptr;
	bool finalized_ = false;
	std::set<std::pair<int, int>> known_pairs_; // (cfg g1, cfg g2) in the order MuJoCo reported them
	std::vector<std::pair<int, int>> pair_list_;
	std::vector<std::pair<int, int>> step_pairs_;
	std::vector<SurfacePluginPtr> plugins, cb_ready_plugins;
	std::vector<double> xpos_, xmat_, vel_, wrench_;
	std::vector<hcs_pair_result> pair_results_;
	std::vector<hcs_face> faces_; // per-face dump of the last step (persistent: 1 << 16 records)
};

// N independent mjData of ONE mjModel stepped together (the north star's "batched independent MuJoCo environments shard by
// environment index across the GPUs"): one hcs_multi context with n_envs = N over the given devices; the host side stays
// C++.  The reference serves one mjData per plugin instance (plugin.cpp:88, 225); this is the same load / collision /
// passive sequence for a whole batch: geoms and contact properties come from the `cs::` custom fields of the model, the
// candidate geom pairs are registered once (setPairs, or recorded from one world's collision pass with recordPair),
// passiveCallback takes all N mjData, makes ONE library call and applies one wrench per geom to every world.
class BatchedContactSurfaces
{
public:
	~BatchedContactSurfaces();
	bool load(const mjModel *m, int n_envs, const std::vector<int> &devices);
	void recordPair(int mujoco_g1, int mujoco_g2); // as collision_cb would see it; ignored for rigid-rigid / unknown geoms
	bool finalize();
	void passiveCallback(const mjModel *m, mjData *const *d, int n);
	hcs_multi *context() { return ctx_; }
	int numPairs() const { return (int)pair_list_.size(); }

private:
	hcs_multi *ctx_ = nullptr;
	int n_envs_     = 0;
	bool applyContactSurfaceForces = true;
	std::map<int, std::shared_ptr<ContactProperties>> contactProperties;
	std::vector<int> cfg_to_mj;
	std::set<std::pair<int, int>> known_pairs_;
	std::vector<std::pair<int, int>> pair_list_;
	std::vector<double> xpos_, xmat_, vel_, wrench_;
	bool finalized_ = false;
};

namespace sensors {

// tactile_sensor_base.h / tactile_sensor_base.cpp:61-120
class TactileSensorBase : public SurfacePlugin
{
public:
	~TactileSensorBase() { delete[] vGeoms; }
	bool load(const mjModel *m, mjData *d) override;
	void update(const mjModel *m, mjData *d, const std::vector<GeomCollisionPtr> &geomCollisions) override;
	void renderCallback(const mjModel *model, mjData *data, mjvScene *scene) override;
	void reset() override;
	bool wantsSensorUpdate(const mjData *d) override;
	// services of the reference: <topic>/set_pause and <topic>/get_state
	void setPause(bool pause);
	std::vector<float> getState(); // blocks until the physics thread has served the request
	const std::vector<float> &lastMessage() const { return tactile_state_values_; }
	const std::string &name() const { return sensorName; }
	int publishCount() const { return publish_count_; }

protected:
	int geomID = -1;
	std::string geomName, topicName, sensorName;
	double updateRate   = 0;
	double updatePeriod = 0;
	double lastUpdate   = -1e300;
	bool visualize      = false;
	mjvGeom *vGeoms     = nullptr;
	int n_vGeom         = 0;
	std::vector<float> tactile_state_values_; // tactile_msgs/TactileState.sensors[0].values
	int publish_count_ = 0;
	virtual void internal_update(const mjModel *, mjData *, const std::vector<GeomCollisionPtr> &) {}
	std::mutex pause_mutex, state_request_mutex;
	std::condition_variable state_cv;
	bool request_state = false;
	bool paused        = false;
};

// flat_tactile_sensor.h / flat_tactile_sensor.cpp:127-214, 262-402
// SENS/config/DynamicFlatTactile.cfg:9-23, the dynamic_reconfigure parameter set of the flat sensor (plain struct:
// ROS is not required to change the parameters at run time)
struct DynamicFlatTactileConfig {
	double update_rate      = 50.0;
	bool visualize          = false;
	bool use_parallel       = true; // kept for interface parity; the GPU path has no serial mode
	double resolution       = 0.025;
	int sampling_resolution = 5;
	int window              = 0; // 0 none, 1 gauss, 2 tukey, 3 square
	double sigma            = -1.0;
};

class FlatTactileSensor : public TactileSensorBase
{
public:
	bool load(const mjModel *m, mjData *d) override;
	// flat_tactile_sensor.cpp:48-125; level == -1 (uint32 max): fetch the current configuration instead of applying one
	void dynamicParamCallback(DynamicFlatTactileConfig &config, uint32_t level, const mjModel *m);
	int cx = 0, cy = 0;

protected:
	void internal_update(const mjModel *m, mjData *d, const std::vector<GeomCollisionPtr> &geomCollisions) override;

private:
	int sensor_index_       = -1;
	int sampling_resolution = 5;
	double resolution       = 0;
	float sigma             = -1.0f;
	int window              = HCS_WINDOW_NONE;
};

// curved_sensor.h / curved_sensor.cpp:111-380 (load), :388-481 (internal_update)
// Config keys as in SENS/config/curved_fingertip.yaml; array values ("taxels", "normals") are whitespace / comma
// separated numbers in this stand-in for XmlRpc.  Surface samples (curved_sensor.cpp:276-283): the reference asks
// vcglib's PoissonSampling for sampleNum = (int)sample_resolution * area points (the cast comes first: 0 for the
// shipped configuration, for which vcglib returns its whole Monte-Carlo pool of 10000 uniform points).  vcglib is
// un-vendored; the adapter restates that behaviour (pool of max(10000, 40 sampleNum) area-weighted uniform points,
// greedy Poisson-disk pruning at radius sqrt(area / (0.7 pi sampleNum)) when sampleNum > 0) with its own fixed-seed
// generator, and exposes the samples for inspection: they are an INPUT of the sensor, not part of the parity claim.
class CurvedSensor : public TactileSensorBase
{
public:
	bool load(const mjModel *m, mjData *d) override;
	const std::vector<double> &samplePoints() const { return sample_pos_; }   // [n][3] geom frame
	const std::vector<double> &sampleNormals() const { return sample_nrm_; }
	const std::vector<double> &taxelPoints() const { return taxel_pos_; }
	const std::vector<double> &taxelNormals() const { return taxel_nrm_; }
	double includeMargin() const { return include_margin; }

protected:
	void internal_update(const mjModel *m, mjData *d, const std::vector<GeomCollisionPtr> &geomCollisions) override;

private:
	int sensor_index_        = -1;
	double include_margin    = 0;
	double sample_resolution = 0;
	std::vector<double> taxel_pos_, taxel_nrm_, sample_pos_, sample_nrm_;
};

// taxel_sensor.h / taxel_sensor.cpp:45-156 (load), :158-478 (internal_update); keys of SENS/config/fingertip.yaml
// and flat_taxel_sensor.yaml.  sample_method "default" only (see hcs_add_taxel_sensor in include/hcs.h).
class TaxelSensor : public TactileSensorBase
{
public:
	bool load(const mjModel *m, mjData *d) override;

protected:
	void internal_update(const mjModel *m, mjData *d, const std::vector<GeomCollisionPtr> &geomCollisions) override;

private:
	int sensor_index_        = -1;
	double include_margin    = 0;
	double sample_resolution = 0;
	double max_pressure      = 0.04; // visualize_max_pressure (taxel_sensor.h:84, taxel_sensor.cpp:70-72)
	std::vector<double> taxel_pos_;
};

} // namespace sensors
} // namespace contact_surfaces
} // namespace mujoco_ros
