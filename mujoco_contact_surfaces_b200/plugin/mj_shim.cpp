// Shim implementations used when MuJoCo is absent (tests / this container only).
// mj_applyFT covers free-jointed bodies: generalized force of a free joint = (force, torque about the
// body's centre of mass) in the world frame, which is what MuJoCo's Jacobian product yields for them.
#ifndef HCS_USE_REAL_MUJOCO
#include "mj_shim.h"

#include <cmath>
#include <cstring>

static int default_collision(const mjModel *, const mjData *, mjContact *, int, int, mjtNum) { return 0; }

mjfCollision mjCOLLISIONFUNC[mjNGEOMTYPES][mjNGEOMTYPES] = {
#define ROW { default_collision, default_collision, default_collision, default_collision, default_collision, default_collision, default_collision, default_collision }
	ROW, ROW, ROW, ROW, ROW, ROW, ROW, ROW
#undef ROW
};

static const char **names_of(const mjModel *m, int type, int &n)
{
	switch (type) {
		case mjOBJ_GEOM: n = m->ngeom; return m->geom_names;
		case mjOBJ_NUMERIC: n = m->nnumeric; return m->numeric_names;
		case mjOBJ_TEXT: n = m->ntext; return m->text_names;
		default: n = 0; return nullptr;
	}
}

int mj_name2id(const mjModel *m, int type, const char *name)
{
	int n;
	const char **names = names_of(m, type, n);
	for (int i = 0; i < n; ++i)
		if (names[i] && std::strcmp(names[i], name) == 0)
			return i;
	return -1;
}

const char *mj_id2name(const mjModel *m, int type, int id)
{
	int n;
	const char **names = names_of(m, type, n);
	return (id >= 0 && id < n) ? names[id] : nullptr;
}

void mj_objectVelocity(const mjModel *, const mjData *d, int, int objid, mjtNum res[6], int)
{
	for (int i = 0; i < 6; ++i)
		res[i] = d->geom_vel6 ? d->geom_vel6[6 * objid + i] : 0.0;
}

void mj_applyFT(const mjModel *m, mjData *d, const mjtNum force[3], const mjtNum torque[3], const mjtNum point[3],
                int body, mjtNum *qfrc_target)
{
	int adr = m->body_dofadr ? m->body_dofadr[body] : -1;
	if (adr < 0)
		return; // static body: no degrees of freedom
	const mjtNum *c = d->xipos + 3 * body;
	mjtNum r[3]     = { point[0] - c[0], point[1] - c[1], point[2] - c[2] };
	mjtNum t[3]     = { r[1] * force[2] - r[2] * force[1], r[2] * force[0] - r[0] * force[2], r[0] * force[1] - r[1] * force[0] };
	for (int i = 0; i < 3; ++i) {
		qfrc_target[adr + i] += force[i];
		qfrc_target[adr + 3 + i] += t[i] + (torque ? torque[i] : 0.0);
	}
}

void mjv_initGeom(mjvGeom *g, int type, const mjtNum size[3], const mjtNum pos[3], const mjtNum mat[9], const float rgba[4])
{
	g->type = type;
	for (int i = 0; i < 3; ++i) {
		g->size[i] = size ? (float)size[i] : 0.1f;
		g->pos[i]  = pos ? (float)pos[i] : 0.f;
	}
	for (int i = 0; i < 9; ++i)
		g->mat[i] = mat ? (float)mat[i] : (i % 4 == 0 ? 1.f : 0.f);
	for (int i = 0; i < 4; ++i)
		g->rgba[i] = rgba ? rgba[i] : 0.5f;
}

void mjv_makeConnector(mjvGeom *g, int type, mjtNum width, mjtNum a0, mjtNum a1, mjtNum a2, mjtNum b0, mjtNum b1, mjtNum b2)
{
	g->type    = type;
	mjtNum d[3] = { b0 - a0, b1 - a1, b2 - a2 };
	mjtNum len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	g->size[0] = g->size[1] = (float)width;
	g->size[2] = (float)(len / 2);
	g->pos[0] = (float)((a0 + b0) / 2), g->pos[1] = (float)((a1 + b1) / 2), g->pos[2] = (float)((a2 + b2) / 2);
}
#endif
