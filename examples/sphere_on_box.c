/* The C ABI from plain C: the reference's sphere_on_box_world.xml (CS/assets/sphere_on_box_world.xml:22-41) as one
 * environment, one step, the reduced wrench per geom that passiveCallback would hand to mj_applyFT.
 *
 *   gcc -std=c99 -I include examples/sphere_on_box.c -L mujoco_contact_surfaces_b200 -lhcs_b200 \
 *       -Wl,-rpath,$PWD/mujoco_contact_surfaces_b200 -o /tmp/sphere_on_box && /tmp/sphere_on_box      (needs a B200)
 */
#include <stdio.h>
#include <string.h>

#include "hcs.h"

#define CHECK(call)                                                                  \
	do {                                                                             \
		int st_ = (call);                                                            \
		if (st_ < 0) {                                                               \
			fprintf(stderr, "%s -> %d: %s\n", #call, st_, hcs_last_error(ctx));      \
			return 1;                                                                \
		}                                                                            \
	} while (0)

int main(void)
{
	hcs_ctx *ctx = NULL;
	hcs_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.device = 0, cfg.n_envs = 1, cfg.representation = HCS_REP_POLYGON, cfg.apply_contact_forces = 1;
	CHECK(hcs_create(&cfg, &ctx));

	/* cs::box0 = [0, 1.0, 0.1, 0.3, 0.3] (rigid), cs::sphere0 = [5e4, 5.0, 0.05, 0.3, 0.3] (soft), in XML order */
	const double box_size[3] = { 0.1, 0.1, 0.1 }, box_props[5] = { 0, 1.0, 0.1, 0.3, 0.3 };
	const double sph_size[3] = { 0.08, 0, 0 }, sph_props[5] = { 5e4, 5.0, 0.05, 0.3, 0.3 };
	int box = hcs_add_geom(ctx, HCS_GEOM_BOX, box_size, NULL, 0, NULL, 0, box_props);
	int sph = hcs_add_geom(ctx, HCS_GEOM_SPHERE, sph_size, NULL, 0, NULL, 0, sph_props);
	CHECK(box);
	CHECK(sph);
	const int32_t g1[1] = { sph }, g2[1] = { box }; /* the pair as MuJoCo hands it to collision_cb */
	CHECK(hcs_set_pairs(ctx, g1, g2, 1));
	CHECK(hcs_finalize(ctx));

	/* poses of one step: mjData.geom_xpos / geom_xmat (row-major), velocities (omega, v); sphere 12 mm into the box top */
	const double xpos[2][3] = { { 0, 0, 0.1 }, { 0.01, -0.02, 0.2 + 0.08 - 0.012 } };
	const double xmat[2][9] = { { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, { 1, 0, 0, 0, 1, 0, 0, 0, 1 } };
	const double vel[2][6]  = { { 0, 0, 0, 0, 0, 0 }, { 0, 0, 0, 0, 0, -0.05 } };
	CHECK(hcs_step(ctx, &xpos[0][0], &xmat[0][0], &vel[0][0], 0));

	double wrench[2][6];
	CHECK(hcs_get_geom_wrenches(ctx, &wrench[0][0]));
	hcs_pair_result pr;
	CHECK(hcs_get_pair_results(ctx, &pr));
	printf("contact surface: %d polygons, area %.6g m^2, centroid (%.4f, %.4f, %.4f)\n", pr.n_polygons, pr.area,
	       pr.centroid[0], pr.centroid[1], pr.centroid[2]);
	printf("sphere: F = (%.4f, %.4f, %.4f) N, torque about the world origin (%.5f, %.5f, %.5f) N m\n", wrench[sph][0],
	       wrench[sph][1], wrench[sph][2], wrench[sph][3], wrench[sph][4], wrench[sph][5]);
	hcs_destroy(ctx);
	return 0;
}
