// Minimal stand-in for the parts of the MuJoCo C API the reference plugin touches
// (inventory: SURVEY.md App. B.5).  MuJoCo is not installed in this environment; the adapter compiles
// against this header so that it can be built and tested here.  Define HCS_USE_REAL_MUJOCO to compile
// the very same adapter sources against <mujoco/mujoco.h>: every name below is MuJoCo's own.
#pragma once

#ifdef HCS_USE_REAL_MUJOCO
#include <mujoco/mujoco.h>
#else
#include <cstddef>

typedef double mjtNum;

enum mjtGeom_ {
	mjGEOM_PLANE = 0, mjGEOM_HFIELD, mjGEOM_SPHERE, mjGEOM_CAPSULE, mjGEOM_ELLIPSOID, mjGEOM_CYLINDER, mjGEOM_BOX,
	mjGEOM_MESH, mjNGEOMTYPES, mjGEOM_ARROW = 100
};
enum mjtObj_ { mjOBJ_BODY = 1, mjOBJ_GEOM = 5, mjOBJ_NUMERIC = 20, mjOBJ_TEXT = 21 };

struct mjModel {
	int ngeom, nbody, nv, nnumeric, ntext, nmesh;
	int *geom_type, *geom_bodyid, *geom_dataid;
	mjtNum *geom_size;
	float *mesh_vert;
	int *mesh_face, *mesh_vertadr, *mesh_vertnum, *mesh_faceadr, *mesh_facenum;
	int *numeric_adr, *numeric_size;
	mjtNum *numeric_data;
	int *text_adr, *text_size;
	char *text_data;
	// names (mj_name2id / mj_id2name)
	const char **geom_names, **numeric_names, **text_names;
	int *body_dofadr; // shim only: first dof of each (free-jointed) body, -1 for static bodies
};

struct mjData {
	mjtNum time;
	mjtNum *geom_xpos, *geom_xmat; // [ngeom*3], [ngeom*9] row-major
	mjtNum *xipos;                 // [nbody*3] body centres of mass (world)
	mjtNum *qfrc_passive;          // [nv]
	mjtNum *geom_vel6;             // shim only: what mj_objectVelocity(mjOBJ_GEOM, id, flg_local=0) returns
};

struct mjContact { mjtNum dist; mjtNum pos[3]; mjtNum frame[9]; int geom1, geom2; };

struct mjvGeom {
	int type;
	float size[3], pos[3], mat[9], rgba[4];
};
struct mjvScene {
	int maxgeom, ngeom;
	mjvGeom *geoms;
};

typedef int (*mjfCollision)(const mjModel *m, const mjData *d, mjContact *con, int g1, int g2, mjtNum margin);
extern mjfCollision mjCOLLISIONFUNC[mjNGEOMTYPES][mjNGEOMTYPES];

int mj_name2id(const mjModel *m, int type, const char *name);
const char *mj_id2name(const mjModel *m, int type, int id);
void mj_objectVelocity(const mjModel *m, const mjData *d, int objtype, int objid, mjtNum res[6], int flg_local);
void mj_applyFT(const mjModel *m, mjData *d, const mjtNum force[3], const mjtNum torque[3], const mjtNum point[3],
                int body, mjtNum *qfrc_target);
void mjv_initGeom(mjvGeom *geom, int type, const mjtNum size[3], const mjtNum pos[3], const mjtNum mat[9],
                  const float rgba[4]);
void mjv_makeConnector(mjvGeom *geom, int type, mjtNum width, mjtNum a0, mjtNum a1, mjtNum a2, mjtNum b0, mjtNum b1,
                       mjtNum b2);
#endif
