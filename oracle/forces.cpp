// CPU ORACLE (test infrastructure, parity unpinned) — the arithmetic the reference owns:
// collision_cb dispatch (plugin.cpp:255-318), evaluateContactSurface (:320-409),
// calcCombined* (:128-159), passiveCallback force law (:411-483), and Drake
// CalcContactFrictionFromSurfaceProperties (multibody/plant/coulomb_friction.cc).
#include "oracle.hpp"

#include <algorithm>

namespace orc {

// plugin.cpp:128-136
static double combined_modulus(double EA, double EB)
{
	if (EA == kInf)
		return EB;
	if (EB == kInf)
		return EA;
	return EA * EB / (EA + EB);
}

// plugin.cpp:138-159
double combined_dissipation(const Geom &a, const Geom &b)
{
	double EA = a.E, EB = b.E, dA = a.dissipation, dB = b.dissipation;
	double Es = combined_modulus(EA, EB);
	if (Es == kInf)
		return 0.5 * (dA + dB);
	double d = 0;
	if (EA != kInf)
		d += Es / EA * dA;
	if (EB != kInf)
		d += Es / EB * dB;
	return d;
}

// coulomb_friction.cc: mu = 2 mu_A mu_B / (mu_A + mu_B), 0 when the denominator is 0
double combined_friction_dynamic(const Geom &a, const Geom &b)
{
	double den = a.mu_d + b.mu_d;
	return den == 0 ? 0.0 : 2 * a.mu_d * b.mu_d / den;
}

// plugin.cpp:320-409
void evaluate_contact_surface(const Scene &sc, PairOut &po)
{
	const Surface &s  = *po.s;
	double dissipation = combined_dissipation(sc.geoms[po.gM], sc.geoms[po.gN]);
	po.pcs.clear();
	for (int face = 0; face < s.num_faces(); ++face) {
		double Ae = s.face_area[face];
		if (!(Ae > 1.0e-14))
			continue;
		V3 nhat   = s.face_normal[face];
		double gM = s.has_gradM ? dot(s.gradM[face], nhat) : kInf;
		double gN = s.has_gradN ? -dot(s.gradN[face], nhat) : kInf;
		if (gM < 1.0e-14 || gN < 1.0e-14)
			continue;
		double g = 1.0 / (1.0 / gM + 1.0 / gN);
		V3 p_WQ  = s.face_centroid[face];
		double p0;
		if (s.tri) {
			const int *idx = &s.face_idx[s.face_first[face]];
			double b       = 1 / 3.;
			p0             = b * s.e[idx[0]];
			p0 += b * s.e[idx[1]];
			p0 += b * s.e[idx[2]];
		} else {
			p0 = dot(s.poly_grad[face], p_WQ) + s.poly_e0[face];
		}
		double fn0 = Ae * p0;
		double k   = Ae * g;
		po.pcs.push_back({ p_WQ, nhat, fn0, k, dissipation, face });
	}
}

// plugin.cpp:411-483 (force law; mj_applyFT is replaced by the per-pair reduced wrench, which is
// exact because mj_applyFT is linear in (force, torque about the application point)).
void passive_forces(const Scene &sc, PairOut &po, const double *xpos, const double *vel)
{
	const double stiction_tolerance = 1.0e-4, relative_tolerance = 1.0e-2;
	double mu = combined_friction_dynamic(sc.geoms[po.gM], sc.geoms[po.gN]);
	int g1 = po.gM, g2 = po.gN;
	V3 xA{ xpos[3 * g1], xpos[3 * g1 + 1], xpos[3 * g1 + 2] }, xB{ xpos[3 * g2], xpos[3 * g2 + 1], xpos[3 * g2 + 2] };
	V3 wA{ vel[6 * g1], vel[6 * g1 + 1], vel[6 * g1 + 2] }, vA{ vel[6 * g1 + 3], vel[6 * g1 + 4], vel[6 * g1 + 5] };
	V3 wB{ vel[6 * g2], vel[6 * g2 + 1], vel[6 * g2 + 2] }, vB{ vel[6 * g2 + 3], vel[6 * g2 + 4], vel[6 * g2 + 5] };
	po.F = po.tau = { 0, 0, 0 };
	po.face_force.clear();
	for (const PointCollision &pc : po.pcs) {
		V3 v_Aq  = vA + cross(wA, pc.p - xA);
		V3 v_Bq  = vB + cross(wB, pc.p - xB);
		V3 v_rel = v_Aq - v_Bq;
		double vn = dot(v_rel, pc.n);
		double fn = std::max(0., 1. - pc.damping * vn) * (pc.fn0 - 0.001 * pc.stiffness * vn);
		V3 f{ 0, 0, 0 };
		if (sc.apply_forces) {
			V3 vt          = v_rel - pc.n * vn;
			double epsilon = stiction_tolerance * relative_tolerance;
			epsilon        = epsilon * epsilon;
			double v_slip  = std::sqrt(norm2(vt) + epsilon);
			V3 that        = vt / v_slip;
			double mu_reg  = mu;
			double s       = v_slip / stiction_tolerance;
			if (s < 1)
				mu_reg = mu * s * (2.0 - s);
			V3 f_slip = -mu_reg * that * fn;
			f         = f_slip + fn * pc.n;
		}
		po.face_force.push_back(f);
		po.F   = po.F + f;
		po.tau = po.tau + cross(pc.p, f);
	}
}

static Xf pose_of(const double *xpos, const double *xmat, int g)
{
	Xf X;
	for (int i = 0; i < 9; ++i)
		X.R.m[i] = xmat[9 * g + i];
	X.p = { xpos[3 * g], xpos[3 * g + 1], xpos[3 * g + 2] };
	return X;
}

// One env step: collision_cb for every configured pair, then passiveCallback.
void step(const Scene &sc, StepState &st, const double *xpos, const double *xmat, const double *vel, bool use_bvh)
{
	int ng = (int)sc.geoms.size();
	st.xpos.assign(xpos, xpos + 3 * ng);
	st.xmat.assign(xmat, xmat + 9 * ng);
	st.vel.assign(vel, vel + 6 * ng);
	st.out.assign(sc.pairs.size(), PairOut());
	st.geom_wrench.assign(ng, { 0, 0, 0, 0, 0, 0 });
	st.n_candidates = 0;
	for (size_t pi = 0; pi < sc.pairs.size(); ++pi) {
		int g1 = sc.pairs[pi][0], g2 = sc.pairs[pi][1];
		const Geom *c1 = &sc.geoms[g1], *c2 = &sc.geoms[g2];
		PairOut &po = st.out[pi];
		if (c1->kind != SOFT && c2->kind != SOFT)
			continue; // rigid-rigid: MuJoCo default collision function (plugin.cpp:270-272)
		std::shared_ptr<Surface> s;
		if (c1->kind == SOFT && c2->kind == SOFT) {
			s = soft_soft(*c1, g1, pose_of(xpos, xmat, g1), *c2, g2, pose_of(xpos, xmat, g2), sc.tri, use_bvh);
		} else {
			if (c1->kind != SOFT) {
				std::swap(c1, c2);
				std::swap(g1, g2);
			}
			Xf p1 = pose_of(xpos, xmat, g1), p2 = pose_of(xpos, xmat, g2);
			if (c2->kind == RIGID_PLANE)
				s = soft_plane(*c1, g1, p1, g2, p2, sc.tri, use_bvh);
			else
				s = soft_rigid(*c1, g1, p1, *c2, g2, p2, sc.tri, use_bvh);
		}
		if (!s)
			continue;
		st.n_candidates += s->n_candidates;
		po.has_surface = true;
		po.s           = s;
		po.gM          = s->gM;
		po.gN          = s->gN;
		evaluate_contact_surface(sc, po);
		passive_forces(sc, po, xpos, vel);
		double A = 0;
		V3 c{ 0, 0, 0 };
		for (int f = 0; f < s->num_faces(); ++f) {
			A += s->face_area[f];
			c = c + s->face_area[f] * s->face_centroid[f];
		}
		po.area     = A;
		po.centroid = A > 0 ? c / A : c;
		double w[6] = { po.F.x, po.F.y, po.F.z, po.tau.x, po.tau.y, po.tau.z };
		for (int i = 0; i < 6; ++i) {
			st.geom_wrench[po.gM][i] += w[i];
			st.geom_wrench[po.gN][i] -= w[i];
		}
	}
}

} // namespace orc
