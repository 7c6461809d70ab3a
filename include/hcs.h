/* hcs.h — C ABI of the B200-native hydroelastic contact-surface engine (libhcs_b200.so).
 *
 * Drop-in boundary for the hot path of ubi-agni/mujoco_contact_surfaces.  Every entry point
 * cites the reference interface it replaces; paths are relative to the reference root,
 *   CS   = mujoco_contact_surfaces,  SENS = mujoco_contact_surface_sensors,
 *   plugin.cpp = CS/src/mujoco_contact_surfaces_plugin.cpp.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; nothing throws across this boundary.
 *   - every function returns HCS_OK (0) or a negative hcs_status; hcs_last_error() explains.
 *   - a context is single-caller (the reference calls the path from the one MuJoCo physics thread,
 *     plugin.cpp:88-95); it owns all device memory and one CUDA stream on one GPU.
 *   - geoms are addressed by their CONFIGURATION INDEX: the order of hcs_add_* calls, which plays the
 *     role of drake_id = GeometryId::get_new_id() (CS/include/mujoco_contact_surfaces/
 *     mujoco_contact_surfaces_plugin.h:175): the geom with the smaller index is M, normals point
 *     out of N into M (plugin.cpp:308-311, 346-348).
 *   - per-step arrays are batched over n_envs independent environments (one env = one mjData of the
 *     reference, plugin.cpp:88); env-major, geom index second:
 *         xpos[env][geom][3]   = mjData.geom_xpos              (plugin.cpp:116-126)
 *         xmat[env][geom][9]   = mjData.geom_xmat, row-major   (plugin.cpp:118-121)
 *         vel [env][geom][6]   = mj_objectVelocity(m,d,mjOBJ_GEOM,id,res,0) = (omega, v)
 *                                                              (plugin.cpp:107-114)
 *   - units SI, world frame, fp64 geometry ("fp64 geometry mode" of the north star).
 */
#ifndef HCS_H_
#define HCS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hcs_ctx hcs_ctx;

typedef enum hcs_status {
	HCS_OK             = 0,
	HCS_E_INVALID      = -1, /* bad argument / call order */
	HCS_E_UNSUPPORTED  = -2, /* geom type the reference does not support either (plugin.cpp:634-647) */
	HCS_E_CUDA         = -3, /* CUDA runtime error (no CPU fallback exists) */
	HCS_E_CAPACITY     = -4, /* candidate / triangle / bin capacity exceeded: results incomplete */
	HCS_E_NOT_FINALIZED = -5
} hcs_status;

/* MuJoCo mjtGeom values used by plugin.cpp:633-808 */
enum { HCS_GEOM_PLANE = 0, HCS_GEOM_HFIELD = 1, HCS_GEOM_SPHERE = 2, HCS_GEOM_CAPSULE = 3, HCS_GEOM_ELLIPSOID = 4,
	   HCS_GEOM_CYLINDER = 5, HCS_GEOM_BOX = 6, HCS_GEOM_MESH = 7 };

/* drake::geometry::HydroelasticContactRepresentation, parsed from the MuJoCo custom text
 * "cs::HydroelasticContactRepresentation" (plugin.cpp:574-591) */
enum { HCS_REP_POLYGON = 0, HCS_REP_TRIANGLE = 1 };

/* FlatTactileSensor windowing (SENS/src/flat_tactile_sensor.cpp:140-168, DynamicFlatTactile.cfg) */
enum { HCS_WINDOW_NONE = 0, HCS_WINDOW_GAUSS = 1, HCS_WINDOW_TUKEY = 2, HCS_WINDOW_SQUARE = 3 };

typedef struct hcs_config {
	int device;                   /* CUDA device ordinal */
	int n_envs;                   /* independent environments resident on this GPU */
	int representation;           /* HCS_REP_* */
	int apply_contact_forces;     /* cs::ApplyContactSurfaceForces (plugin.cpp:603-611) */
	int max_candidates_per_slice; /* 0 = automatic; >0: each pair's candidate pool holds this many candidates per
	                               * (env, query slice) unit ON AVERAGE (the pool is shared by the whole batch) */
	int max_faces;                /* >0: keep a per-face dump (PointCollision views) of that capacity */
	int max_tactile_triangles;    /* 0 = automatic; triangle pool of the tactile stage, whole batch */
	int max_triangles_per_taxel;  /* 0 = automatic (32); AVERAGE (triangle, taxel) overlaps per taxel the bins hold */
	void *stream;                 /* cudaStream_t to run on, NULL = context-owned stream */
	int face_vertices;            /* != 0 (needs max_faces > 0): the per-face dump also keeps every face's vertices
	                               * (what visualizeMeshElement draws, plugin.cpp:525-555); see hcs_get_face_vertices */
} hcs_config;

/* result of one geom pair in one env: what passiveCallback applies (plugin.cpp:411-483), reduced.
 * F acts on geom gM at the world origin-referenced torque tau (sum of p x f); -F, -tau act on gN. */
typedef struct hcs_pair_result {
	double F[3];
	double tau[3];
	double centroid[3]; /* area-weighted centroid of the contact surface (ContactSurface::centroid) */
	double area;        /* ContactSurface::total_area */
	int32_t gM, gN;     /* configuration indices, gM < gN */
	int32_t n_polygons; /* contact polygons emitted (= |emitted candidate set|) */
	int32_t n_faces;    /* faces of the surface (kPolygon: = n_polygons; kTriangle: fan triangles) */
	int32_t n_points;   /* PointCollisions that pass plugin.cpp:345 and :362 */
	int32_t n_candidates; /* pair-evals spent on this pair: LBVH leaf hits / tets sliced (cull + clip) */
	int32_t n_clipped;    /* pair-evals that survived the early-outs and ran the clipper */
	int32_t reserved;
} hcs_pair_result;

/* one PointCollision (CS/include/mujoco_contact_surfaces/common_types.h:48-56) plus provenance */
typedef struct hcs_face {
	double p[3], n[3];
	double fn0, stiffness, damping;
	double f[3];              /* force applied to gM at p (plugin.cpp:473-475) */
	int32_t env, pair;
	int32_t elemM, elemN;     /* mesh elements that produced the polygon */
	int32_t nverts, face;     /* polygon vertex count; face index inside the polygon's fan */
} hcs_face;

/* --- lifetime ------------------------------------------------------------------------------- */
/* replaces MujocoContactSurfacesPlugin::load() up to parseMujocoCustomFields (plugin.cpp:208-222) */
int hcs_create(const hcs_config *cfg, hcs_ctx **out);
/* replaces ~MujocoContactSurfacesPlugin (plugin.cpp:191-206) */
void hcs_destroy(hcs_ctx *ctx);
const char *hcs_last_error(const hcs_ctx *ctx); /* ctx may be NULL: error of the last failed create */

/* --- configuration: one call per `cs::<geom>` numeric, in XML order ------------------------------ */
/* replaces the geom switch of parseMujocoCustomFields (plugin.cpp:613-812).
 *   size  = mjModel.geom_size[3*id..]  (sphere r | ellipsoid semi-axes | cylinder r,half-length | box half sizes)
 *   props = {hydroelasticModulus, dissipation, resolutionHint, staticFriction, dynamicFriction};
 *           modulus > 0 => SOFT (tet mesh + linear pressure field), else RIGID (triangle surface / plane)
 *   mesh_vert/mesh_face: mjModel.mesh_vert (float32) / mesh_face (int32) of a MESH geom (plugin.cpp:745-763)
 * returns the configuration index >= 0, HCS_E_UNSUPPORTED for plane-soft / hfield / capsule. */
int hcs_add_geom(hcs_ctx *ctx, int mj_geom_type, const double size[3], const float *mesh_vert, int n_vert,
                 const int32_t *mesh_face, int n_face, const double props[5]);
/* user-supplied meshes (same role as the drake::geometry::VolumeMesh / TriangleSurfaceMesh a
 * ContactProperties holds, mujoco_contact_surfaces_plugin.h:147-150) */
int hcs_add_soft_mesh(hcs_ctx *ctx, const double *verts, int n_vert, const int32_t *tets, int n_tet,
                      const double *vertex_pressure, const double props[5]);
int hcs_add_rigid_mesh(hcs_ctx *ctx, const double *verts, int n_vert, const int32_t *tris, int n_tri,
                       const double props[5]);
/* replaces MujocoContactSurfacesPlugin::onGeomChanged (plugin.cpp:828-975): rebuild one geom after a
 * size change (the reference's Q2 bugs are not reproduced: the new pressure field IS installed). */
int hcs_update_geom(hcs_ctx *ctx, int geom, const double size[3]);

/* geom pairs MuJoCo's collision pass hands to collision_cb (plugin.cpp:255-318), same for every env.
 * (g1,g2) in the order MuJoCo would pass them; rigid-rigid pairs are accepted and ignored. */
/* Per-environment sizes of one geom (domain randomisation; SURVEY.md section 8 f3 "GPU mesh/field/LBVH (re)build ... for
 * domain-randomised sizes per env"; the reference rebuilds ONE geom of its one mjData in onGeomChanged, plugin.cpp:828-975).
 * sizes = [n_envs][3] in the convention of hcs_add_geom.  Every environment gets the mesh, pressure field and LBVH the
 * single-size path would build for its size; all environments must lead to the SAME topology (sphere / ellipsoid: same
 * refinement level, box grid: same cell counts, ...), else HCS_E_UNSUPPORTED with the first offending environment in the
 * error text.  Sphere and ellipsoid vertices and pressures are generated on the GPU from the unit mesh (bit-identical to
 * the host generator); other shapes are generated per environment on the host; fields, element records and the LBVH
 * are built on the GPU for every environment.  sizes == NULL returns the geom to one size for all environments.
 * Before or after hcs_finalize (then the context is rebuilt, like hcs_update_geom). */
int hcs_set_env_sizes(hcs_ctx *ctx, int geom, const double *sizes);
int hcs_set_pairs(hcs_ctx *ctx, const int32_t *g1, const int32_t *g2, int n_pairs);

/* replaces FlatTactileSensor::load (SENS/src/flat_tactile_sensor.cpp:127-214): taxel grid
 * cx = floor(2*size[0]/resolution + 0.1), cy likewise, sampling_resolution^2 rays per taxel.
 * geom: a geom already added whose three size entries are positive (box, ellipsoid: the reference reads
 * geom_size[0..2] of whatever geom carries the sensor, :192-197); returns the sensor index. */
int hcs_add_flat_sensor(hcs_ctx *ctx, int geom, double resolution, int sampling_resolution, int window, float sigma);
int hcs_sensor_dims(const hcs_ctx *ctx, int sensor, int *cx, int *cy);
/* replaces the part of FlatTactileSensor::dynamicParamCallback (flat_tactile_sensor.cpp:48-125, the dynamic_reconfigure
 * set of SENS/config/DynamicFlatTactile.cfg) that changes the computation: sampling_resolution, window, sigma (the
 * resolution is fixed after load in the reference too, :66; update_rate / visualize / use_parallel are host-side
 * state of the adapter).  The sensor's ray grid and window table are rebuilt; a finalized context stays finalized. */
int hcs_update_flat_sensor(hcs_ctx *ctx, int sensor, int sampling_resolution, int window, float sigma);

/* replaces CurvedSensor (SENS/src/curved_sensor.cpp): load() :111-380 and internal_update() :388-481.
 * taxel_pos / taxel_nrm: [n_taxels][3] in the sensor geom's frame (taxel_nrm may be NULL: no 45-degree test);
 * sample_pos / sample_nrm: [n_samples][3] surface sample points with unit normals in the same frame.  The reference
 * draws the samples with vcglib's Poisson-disk sampler (:271-283, un-vendored); here the caller supplies them, and
 * the assignment of samples to taxels (distance < include_margin, normal within 45 degrees) and the weights
 * (include_margin - distance)^2 are computed as in :337-360.  Every step (with_sensors != 0) each assigned sample
 * casts one float32 ray from its world position along the inward normal; the nearest contact-surface triangle hit
 * with 0 < t < include_margin contributes weight * e_MN(hit) to the taxel.  Returns the curved sensor index. */
int hcs_add_curved_sensor(hcs_ctx *ctx, int geom, int n_taxels, const double *taxel_pos, const double *taxel_nrm,
                          int n_samples, const double *sample_pos, const double *sample_nrm, double include_margin);
int hcs_curved_sensor_info(const hcs_ctx *ctx, int sensor, int *n_taxels, int *n_rays, int *n_assignments);
/* out: float [n_envs][n_taxels], TactileState.sensors[0].values of each environment (:479) */
int hcs_get_curved_values(hcs_ctx *ctx, int sensor, float *out);
const float *hcs_device_curved_values(hcs_ctx *ctx, int sensor);

/* replaces TaxelSensor (SENS/src/taxel_sensor.cpp): load() :45-156, internal_update() :158-478.
 * taxel_pos: [n_taxels][3] in the sensor geom's frame.  method: 0 closest, 1 weighted, 2 mean, 3 squared (:79-91).
 * sample_method 0 = "default": every step (with_sensors != 0) each contact-surface triangle of the sensor geom is
 * sampled on the barycentric lattice of :191-211.  sample_method 1 = "area_importance" (:211-254, the method the
 * reference's fingertip.yaml uses): one sample per stratum of sample_resolution * total_area along the cumulative
 * triangle area of every surface, uniform in the owning triangle, drawn from one std::default_random_engine
 * (minstd_rand0, default seed, libstdc++ generate_canonical) that restarts every update.  Which random numbers a
 * triangle gets depends on the triangle order; Drake's is not observable through the reference, ours is canonical
 * (pairs in pair order, polygons by (elemM, elemN), fan triangles in fan order), so values agree with the reference
 * in distribution, and with this repo's oracle to 1e-6.  Limits of this method: at most 4096 contact-surface triangles
 * of the sensor geom per environment (they are ordered in shared memory), HCS_E_CAPACITY beyond.
 * Every taxel is evaluated from the samples within include_margin.  The reference's observable behaviour is
 * reproduced, including its quirks: weighted and mean fall through to squared (value =
 * sample_resolution * sum (include_margin - d)^2 |p|), closest keeps the pressure only when `visualize` is on (else 0),
 * taxels without a sample in range keep the value of the previous update, an update without any sample zeroes the
 * message.  Returns the taxel sensor index. */
int hcs_add_taxel_sensor(hcs_ctx *ctx, int geom, int n_taxels, const double *taxel_pos, double include_margin,
                         double sample_resolution, int method, int visualize, int sample_method);
/* out: float [n_envs][n_taxels] */
int hcs_get_taxel_values(hcs_ctx *ctx, int sensor, float *out);
const float *hcs_device_taxel_values(hcs_ctx *ctx, int sensor);

/* builds fields, tet half spaces and LBVHs on the GPU, allocates per-env buffers */
int hcs_finalize(hcs_ctx *ctx);

/* --- per step ----------------------------------------------------------------------------------- */
/* replaces, for all envs at once: every collision_cb (plugin.cpp:255-318) + evaluateContactSurface
 * (:320-409) + the force loop of passiveCallback (:411-483) + FlatTactileSensor::bvh_update
 * (SENS/src/flat_tactile_sensor.cpp:262-402) when sensors exist and with_sensors != 0.
 * HOST pointers; copies in/out are part of the call (this is the end-to-end entry point): poses and velocities
 * go to the device, the per-geom wrenches (what passiveCallback applies) and, with with_sensors, the sensor
 * outputs come back before the call returns.  The per-pair results are diagnostics without a counterpart in the
 * reference: they stay on the device until hcs_get_pair_results / hcs_get_counters / hcs_fetch_results ask. */
int hcs_step(hcs_ctx *ctx, const double *xpos, const double *xmat, const double *vel, int with_sensors);
/* same with DEVICE pointers; asynchronous on the context stream, results stay on the GPU */
int hcs_step_device(hcs_ctx *ctx, const double *d_xpos, const double *d_xmat, const double *d_vel,
                    int with_sensors);
/* Pipelined end-to-end step (SURVEY.md section 8b "Threading": the reference's physics thread calls the path once per
 * mj_step; a batched caller overlaps its own work, and the copies, with the GPU).  HOST pointers as hcs_step; the call
 * returns as soon as the work is queued: the host-to-device copies of this step's inputs run on a copy stream and
 * overlap the kernels of the previous step, the results go to the CALLER-OWNED host buffers of `out` (any member may be
 * NULL; pinned memory keeps the copies asynchronous) on a second copy stream while the next step's kernels run.  Two
 * steps may be in flight; a third call first finishes the oldest one.  Inputs may be reused as soon as the call
 * returns only if they are pageable (the copy is staged); pinned inputs must stay untouched until hcs_wait(ticket) of
 * that step or the next hcs_step_async call but one.  Geometry, pairs and sensors are those of hcs_step. */
typedef struct hcs_outputs {
	double *geom_wrench;           /* [n_envs * n_geoms * 6], as hcs_get_geom_wrenches */
	float **sensor_images;         /* [n flat sensors] -> [n_envs * cx * cy], as hcs_get_sensor_image (with_sensors) */
	float **curved_values;         /* [n curved sensors] -> [n_envs * n_taxels] */
	float **taxel_values;          /* [n taxel sensors] -> [n_envs * n_taxels] */
	hcs_pair_result *pair_results; /* [n_envs * n_pairs] diagnostics, as hcs_get_pair_results */
} hcs_outputs;
int hcs_step_async(hcs_ctx *ctx, const double *xpos, const double *xmat, const double *vel, int with_sensors,
                   const hcs_outputs *out, int64_t *ticket);
/* blocks until the step's results are in the caller's buffers; HCS_OK or the step's capacity / range error */
int hcs_wait(hcs_ctx *ctx, int64_t ticket);
/* wait for the stream, check the capacity flags (HCS_E_CAPACITY) */
int hcs_sync(hcs_ctx *ctx);
/* copy the device results of the last hcs_step_device into the context's pinned host mirrors */
int hcs_fetch_results(hcs_ctx *ctx, int with_sensors);

/* --- results (valid until the next step) ------------------------------------------------------------ */
int hcs_n_geoms(const hcs_ctx *ctx);
int hcs_n_pairs(const hcs_ctx *ctx);
/* out[n_envs * n_pairs], env-major */
int hcs_get_pair_results(hcs_ctx *ctx, hcs_pair_result *out);
/* out[n_envs * n_geoms * 6]: (F, tau about the world origin) the contact surfaces apply to each geom;
 * the adapter feeds this to ONE mj_applyFT per body instead of two per face (plugin.cpp:477-482) */
int hcs_get_geom_wrenches(hcs_ctx *ctx, double *out);
/* out[n_envs * cx * cy], index x + cy*y inside an image (SENS/src/flat_tactile_sensor.cpp:396-397) */
int hcs_get_sensor_image(hcs_ctx *ctx, int sensor, float *out);
/* device-resident views of the same results (no copy) */
const hcs_pair_result *hcs_device_pair_results(hcs_ctx *ctx);
const double *hcs_device_geom_wrenches(hcs_ctx *ctx);
const float *hcs_device_sensor_image(hcs_ctx *ctx, int sensor);

/* per-face dump = the GeomCollision/PointCollision views handed to SurfacePlugin::update
 * (CS/include/mujoco_contact_surfaces/plugin_utils.h:94); needs cfg.max_faces > 0.
 * returns the number of faces written (<= cap) or a negative status. Order is unspecified. */
int hcs_get_faces(hcs_ctx *ctx, hcs_face *out, int cap);
/* vertices of the dumped faces (world frame), the element of ContactSurface::poly_mesh_W() / tri_mesh_W() that
 * visualizeMeshElement(pc.face, mesh, fn) walks (plugin.cpp:509-516, 525-555); needs cfg.face_vertices != 0.
 * out[i*24 .. i*24+3*nv) = the nv vertices of face i of hcs_get_faces (same order of faces, same step), wound
 * counter-clockwise about the face normal hcs_face.n; nv = hcs_face.nverts (kPolygon, <= 8) or 3 (kTriangle:
 * TriMeshBuilder's fan triangle (previous, next, centroid)).  Returns the number of faces (as hcs_get_faces). */
#define HCS_FACE_VERTEX_STRIDE 24
int hcs_get_face_vertices(hcs_ctx *ctx, double *out, int cap);
/* emitted candidate set of one pair in one env: triples (elemM, elemN, nverts); returns count */
int hcs_get_emitted(hcs_ctx *ctx, int env, int pair, int32_t *out, int cap);
/* kTriangle contact-surface triangle soup of one env (world frame), 12 doubles per triangle:
 * 9 vertex coordinates + 3 vertex pressures; only pairs touching a sensor geom are kept. */
int hcs_get_tactile_triangles(hcs_ctx *ctx, int env, double *out, int cap);
/* the pair index of every triangle of hcs_get_tactile_triangles(env), same order (the soup is sorted by pair first):
 * lets a host adapter hand each ContactSurface its own triangles (reference: one ContactSurface per geom pair,
 * mujoco_contact_surfaces_plugin.cpp:301-315); returns count */
int hcs_get_tactile_triangle_pairs(hcs_ctx *ctx, int env, int32_t *pair_out, int cap);

/* --- introspection (tests, DESIGN.md evidence) ---------------------------------------------------------- */
/* info = {kind (0 rigid mesh, 1 soft, 2 plane), n_vertices, n_elements} */
int hcs_geom_info(const hcs_ctx *ctx, int geom, int info[3]);
/* copies the device-resident mesh back: verts[nv*3], elems[ne*(4|3)], and for soft geoms
 * pressure[nv], grad[ne*3], e0[ne]; for rigid geoms grad receives the unit face normals */
int hcs_get_mesh(hcs_ctx *ctx, int geom, double *verts, int32_t *elems, double *pressure, double *grad, double *e0);
/* LBVH of a soft geom as resident on the GPU: 64-byte records {float llo[3], lhi[3], rlo[3], rhi[3]; int32 left,
 * right; float pad[2]} (child >= 0: internal node, < 0: tet ~child); returns the node count (n_tets - 1, min 1) */
int hcs_get_lbvh(hcs_ctx *ctx, int geom, void *out_nodes, int max_nodes);
/* counters of the last step: {candidate pair-evals, polygons, faces, tactile triangles, kernels launched} */
int hcs_get_counters(hcs_ctx *ctx, int64_t out[5]);
/* per-stage GPU time of the last step in ms (CUDA events on the context stream); needs hcs_set_profiling(1):
 * {setup, broadphase, narrowphase, reduce, tactile_bin, tactile_raster, total} */
int hcs_set_profiling(hcs_ctx *ctx, int enable);
int hcs_get_stage_ms(hcs_ctx *ctx, float out[7]);
const char *hcs_version(void);

/* --- one context spanning several GPUs ---------------------------------------------------------------------------
 * North star: "batched independent MuJoCo environments shard by environment index across the GPUs of one box, with no
 * NCCL on the step path"; SURVEY.md section 8(b) "Threading".  A multi-device context owns one single-device context per
 * entry of devices[] (a device may appear more than once), each with a contiguous block of the n_envs environments
 * (sizes differ by at most one: block k starts at k * (n / D) + min(k, n % D)), and one host thread per block: every
 * configuration call is replayed on all blocks (geometry is replicated at finalize), a step hands every block its slice
 * of the env-major host arrays and runs the blocks concurrently, results land in the caller's env-major arrays.  The
 * host code stays C++ (std::thread); nothing is exchanged between the GPUs.  cfg->device and cfg->stream are ignored
 * (every block owns its stream), cfg->n_envs is the TOTAL.  Results are bit-identical to a single context with the same
 * n_envs: environments never influence each other (exact integer accumulators, hcs_internal.h). */
typedef struct hcs_multi hcs_multi;
int hcs_multi_create(const hcs_config *cfg, const int *devices, int n_devices, hcs_multi **out);
void hcs_multi_destroy(hcs_multi *m);
const char *hcs_multi_last_error(const hcs_multi *m);
int hcs_multi_n_blocks(const hcs_multi *m);
/* first environment and environment count of block k; the block's own context (for everything not mirrored here) */
int hcs_multi_block(const hcs_multi *m, int k, int *env_start, int *env_count, hcs_ctx **ctx);
int hcs_multi_add_geom(hcs_multi *m, int mj_geom_type, const double size[3], const float *mesh_vert, int n_vert,
                       const int32_t *mesh_face, int n_face, const double props[5]);
int hcs_multi_add_soft_mesh(hcs_multi *m, const double *verts, int n_vert, const int32_t *tets, int n_tet,
                            const double *vertex_pressure, const double props[5]);
int hcs_multi_add_rigid_mesh(hcs_multi *m, const double *verts, int n_vert, const int32_t *tris, int n_tri,
                             const double props[5]);
int hcs_multi_update_geom(hcs_multi *m, int geom, const double size[3]);
int hcs_multi_set_env_sizes(hcs_multi *m, int geom, const double *sizes); /* [n_envs][3] of the whole batch */
int hcs_multi_set_pairs(hcs_multi *m, const int32_t *g1, const int32_t *g2, int n_pairs);
int hcs_multi_add_flat_sensor(hcs_multi *m, int geom, double resolution, int sampling_resolution, int window, float sigma);
int hcs_multi_finalize(hcs_multi *m);
/* HOST arrays of the whole batch, as hcs_step; returns when every block has finished (first error wins) */
int hcs_multi_step(hcs_multi *m, const double *xpos, const double *xmat, const double *vel, int with_sensors);
/* the same through every block's pipelined entry point (hcs_step_async): only hands the step to the blocks' host threads
 * and returns (they queue their slices on their own time; errors surface in hcs_multi_wait); out members are env-major
 * arrays of the whole batch; the input arrays must stay untouched until hcs_multi_wait(ticket); at most two steps may be
 * un-waited */
int hcs_multi_step_async(hcs_multi *m, const double *xpos, const double *xmat, const double *vel, int with_sensors,
                         const hcs_outputs *out, int64_t *ticket);
int hcs_multi_wait(hcs_multi *m, int64_t ticket);
int hcs_multi_get_geom_wrenches(hcs_multi *m, double *out);        /* [n_envs][n_geoms][6] */
int hcs_multi_get_pair_results(hcs_multi *m, hcs_pair_result *out); /* [n_envs][n_pairs] */
int hcs_multi_get_sensor_image(hcs_multi *m, int sensor, float *out); /* [n_envs][cx*cy] */

#ifdef __cplusplus
}
#endif
#endif /* HCS_H_ */
