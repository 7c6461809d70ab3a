"""Known-answer and invariance tests of the CPU oracle (oracle/).

The reference has no tests, golden vectors or runnable build in this environment, so
the oracle is pinned against hand-derivable answers instead: "parity unpinned" with respect to reference
outputs, as stated in oracle/oracle.hpp and DESIGN.md.
"""
import numpy as np
import pytest

from oracle.oracle import (GEOM_BOX, GEOM_CYLINDER, GEOM_ELLIPSOID, GEOM_MESH, GEOM_PLANE, GEOM_SPHERE, KIND_SOFT,
                           OracleScene)

I3 = np.eye(3).reshape(-1)
SOFT = [5e4, 5.0, 0.05, 0.3, 0.3]
RIGID = [0, 1.0, 0.1, 0.3, 0.3]


def tet_volumes(m):
    v, t = m["verts"], m["elems"]
    a, b, c, d = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]], v[t[:, 3]]
    return np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a) / 6


def rot(axis, ang):
    axis = np.asarray(axis, dtype=float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


# ---- stage 1: meshes and pressure fields ---------------------------------------------------------------
def test_sphere_mesh_counts_match_reference_config():
    """sphere_on_box_world.xml: r=0.08, hint 0.05 -> level 2 -> 128 tets / 67 vertices."""
    s = OracleScene()
    g = s.add_geom(GEOM_SPHERE, [0.08], SOFT)
    m = s.geom_mesh(g)
    assert m["elems"].shape == (128, 4) and m["verts"].shape == (67, 3)
    vol = tet_volumes(m)
    assert (vol > 0).all()
    assert np.allclose(np.linalg.norm(m["verts"][1:], axis=1), 0.08, rtol=1e-15)
    # inscribed polyhedron: volume below, and converging to, the ball
    assert 0.8 * 4 / 3 * np.pi * 0.08 ** 3 < vol.sum() < 4 / 3 * np.pi * 0.08 ** 3
    assert m["pressure"][0] == 5e4 and np.all(m["pressure"][1:] == 0)


@pytest.mark.parametrize("r,hint,level", [(0.08, 0.05, 2), (0.05, 0.015, 3), (0.08, 1.0, 0), (1.0, 0.02, 7)])
def test_sphere_refinement_level(r, hint, level):
    s = OracleScene()
    m = s.geom_mesh(s.add_geom(GEOM_SPHERE, [r], [1e5, 1, hint, 0, 0]))
    assert len(m["elems"]) == 8 * 4 ** level


def test_rigid_box_surface_counts():
    """box 0.2^3, hint 0.1 -> 48 triangles / 26 vertices; hint 0.05 -> 192 triangles."""
    s = OracleScene()
    m = s.geom_mesh(s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], RIGID))
    assert m["elems"].shape == (48, 3) and m["verts"].shape == (26, 3)
    m2 = s.geom_mesh(s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1, 0.05, 0.3, 0.3]))
    assert m2["elems"].shape == (192, 3)
    # outward normals, closed surface: sum of area vectors is zero, enclosed volume is the box volume
    v, t = m["verts"], m["elems"]
    a, b, c = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    assert np.allclose(np.cross(b - a, c - a).sum(0), 0, atol=1e-15)
    assert np.isclose(np.einsum("ij,ij->i", np.cross(a, b), c).sum() / 6, 0.008)
    assert np.allclose(np.einsum("ij,ij->i", m["normal"], (a + b + c) / 3) > 0, True)


@pytest.mark.parametrize("half", [[0.2, 0.2, 0.02], [0.1, 0.1, 0.1], [0.05, 0.04, 0.03], [0.3, 0.1, 0.1]])
def test_box_medial_axis_mesh(half):
    """MA box: positive tets that tile the box; the pressure field is the exact distance field."""
    s = OracleScene()
    m = s.geom_mesh(s.add_geom(GEOM_BOX, half, [5e4, 5, 0, 0.3, 0.3]))
    vol = tet_volumes(m)
    assert (vol > 0).all()
    assert np.isclose(vol.sum(), 8 * np.prod(half), rtol=1e-13)
    assert len(m["verts"]) <= 16
    # linear interpolation of the vertex pressures reproduces E * dist / min_half at random interior points
    rng = np.random.default_rng(0)
    h = np.asarray(half)
    for p in rng.uniform(-1, 1, size=(200, 3)) * h:
        expect = 5e4 * np.min(h - np.abs(p)) / h.min()
        vals = m["grad"] @ p + m["e0"]
        # find the containing tet through barycentric coordinates
        inside = []
        for k, t in enumerate(m["elems"]):
            A = np.c_[m["verts"][t].T, np.ones(4)].T if False else np.vstack([m["verts"][t].T, np.ones(4)])
            bary = np.linalg.solve(A, np.r_[p, 1.0])
            if (bary > -1e-9).all():
                inside.append(k)
        assert inside, "point not covered by the mesh"
        assert np.allclose(vals[inside], expect, rtol=1e-9, atol=1e-9)


def test_foam_box_is_reference_size():
    s = OracleScene()
    m = s.geom_mesh(s.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5, 0, 0.3, 0.3]))
    assert m["verts"].shape == (12, 3) and 20 <= len(m["elems"]) <= 24


@pytest.mark.parametrize("r,hl", [(0.05, 0.1), (0.05, 0.05), (0.1, 0.03)])
def test_cylinder_medial_axis_mesh(r, hl):
    s = OracleScene()
    hint = 0.02
    m = s.geom_mesh(s.add_geom(GEOM_CYLINDER, [r, hl, 0], [5e4, 5, hint, 0.3, 0.3]))
    vol = tet_volumes(m)
    n = max(3, int(np.ceil(2 * np.pi * r / hint)))
    prism = 0.5 * n * r * r * np.sin(2 * np.pi / n) * 2 * hl
    assert (vol > 0).all() and np.isclose(vol.sum(), prism, rtol=1e-12)
    assert m["pressure"].max() == 5e4 and m["pressure"].min() == 0


def test_ellipsoid_and_convex_mesh():
    s = OracleScene()
    m = s.geom_mesh(s.add_geom(GEOM_ELLIPSOID, [0.05, 0.04, 0.03], [5e4, 5, 0.01, 0.3, 0.3]))
    assert (tet_volumes(m) > 0).all()
    q = m["verts"][1:] / [0.05, 0.04, 0.03]
    assert np.allclose(np.linalg.norm(q, axis=1), 1, rtol=1e-14)
    # convex "centroid fan" of a cube given as a triangle mesh (plugin.cpp:161-187, 767-787)
    cube = s.geom_mesh(s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], RIGID))
    verts = (cube["verts"] + [0.3, 0.2, 0.1]).astype(np.float32)
    c = s.geom_mesh(s.add_geom(GEOM_MESH, [0, 0, 0], [1e5, 1, 0, 0.3, 0.3], verts, cube["elems"]))
    assert c["kind"] == KIND_SOFT and len(c["elems"]) == 48 and len(c["verts"]) == 27
    assert np.allclose(c["verts"][-1], [0.3, 0.2, 0.1], atol=1e-7)  # centroid of the enclosed volume
    assert c["pressure"][-1] == 1e5 and np.all(c["pressure"][:-1] == 0)
    assert (tet_volumes(c) > 0).all()


def test_unsupported_geoms_are_rejected_like_the_reference():
    s = OracleScene()
    with pytest.raises(ValueError):
        s.add_geom(GEOM_PLANE, [0, 0, 1], SOFT)  # soft plane (plugin.cpp:635-636)
    with pytest.raises(ValueError):
        s.add_geom(3, [0.1, 0.1, 0], SOFT)  # capsule (plugin.cpp:645-647)
    with pytest.raises(ValueError):
        s.add_geom(1, [0.1, 0.1, 0], RIGID)  # hfield


# ---- stage 3: the three queries -------------------------------------------------------------------------
UNIT_TET = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])


def unit_tet_scene(tri_verts, pressure=(0, 0, 0, 1.0), triangle=False):
    s = OracleScene(triangle_representation=triangle)
    soft = s.add_raw_soft(UNIT_TET, [[0, 1, 2, 3]], pressure, [1.0, 0, 0, 0, 0])
    rigid = s.add_raw_rigid(tri_verts, [[0, 1, 2]], [0, 1, 0, 0, 0])
    s.set_pairs([[soft, rigid]])
    return s


def test_tet_triangle_clip_known_polygon():
    """Big triangle in the plane z = 0.25 (normal -z, along -grad within 5pi/8? no: normal must satisfy the cull)
    clipped by the unit tet gives the triangle x,y>=0, x+y<=0.75 with area 0.28125."""
    # pressure p = z (gradient +z); triangle normal +z passes the cull (cos = 1)
    tri = np.array([[-5, -5, 0.25], [5, -5, 0.25], [0, 8, 0.25]])
    for bvh in (False, True):
        s = unit_tet_scene(tri)
        s.step(np.zeros((2, 3)), np.stack([I3, I3]), use_bvh=bvh)
        r = s.pair_result(0)
        assert r["n_polygons"] == 1 and s.pair_emitted(0).tolist() == [[0, 0, 3]]
        assert np.isclose(r["area"], 0.5 * 0.75 ** 2, rtol=1e-14)
        assert np.allclose(r["centroid"], [0.25, 0.25, 0.25], rtol=1e-14)
        # fn0 = area * p(centroid) = 0.28125 * 0.25, pushing the soft tet (M, lower id) along +z
        assert np.allclose(r["F"], [0, 0, 0.28125 * 0.25], rtol=1e-13, atol=1e-16)


def test_tet_triangle_normal_cull():
    tri = np.array([[-5, -5, 0.25], [0, 8, 0.25], [5, -5, 0.25]])  # normal -z: cos = -1 < cos(5pi/8)
    s = unit_tet_scene(tri)
    s.step(np.zeros((2, 3)), np.stack([I3, I3]), use_bvh=False)
    assert not s.pair_result(0)["has_surface"]


def test_tet_triangle_kTriangle_fan_conserves_area_and_force():
    tri = np.array([[-5, -5, 0.25], [5, -5, 0.25], [0, 8, 0.25]])
    sp, st = unit_tet_scene(tri), unit_tet_scene(tri, triangle=True)
    for s in (sp, st):
        s.step(np.zeros((2, 3)), np.stack([I3, I3]), use_bvh=False)
    rp, rt = sp.pair_result(0), st.pair_result(0)
    assert rt["n_faces"] == 3 and rp["n_faces"] == 1
    assert np.isclose(rt["area"], rp["area"], rtol=1e-14)
    assert np.allclose(rt["F"], rp["F"], rtol=1e-13)  # centroid quadrature is exact for a linear field
    tris = st.pair_triangles(0)
    assert tris.shape == (3, 12) and np.allclose(tris[:, 6:9], [0.25, 0.25, 0.25])


def test_half_space_slices_all_fourteen_marching_tet_codes():
    """Slice the unit tet with a plane for each sign pattern; the polygon area must equal the analytic
    cross-section (computed by sampling-free convex-hull area) and its normal the plane normal."""
    rng = np.random.default_rng(1)
    seen = set()
    for trial in range(400):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        d = rng.uniform(-0.2, 0.8)
        h = UNIT_TET @ n - d
        code = sum(1 << i for i in range(4) if h[i] > 0)
        if code in (0, 15) or np.abs(h).min() < 1e-3:
            continue
        seen.add(code)
        # plane geom: z axis = n, origin on the plane
        z = n
        x = np.cross(z, [1, 0, 0] if abs(z[0]) < 0.9 else [0, 1, 0])
        x /= np.linalg.norm(x)
        R = np.stack([x, np.cross(z, x), z], axis=1)
        s = OracleScene()
        soft = s.add_raw_soft(UNIT_TET, [[0, 1, 2, 3]], [1.0, 2.0, 3.0, 4.0], [1.0, 0, 0, 0, 0])
        plane = s.add_geom(GEOM_PLANE, [0, 0, 1], [0, 1, 0, 0, 0])
        s.set_pairs([[plane, soft]])
        s.step(np.array([np.zeros(3), n * d]), np.stack([I3, R.reshape(-1)]), use_bvh=False)
        r = s.pair_result(0)
        # analytic: intersection points of the 6 edges with the plane
        pts = []
        for a, b in [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]:
            if (h[a] > 0) != (h[b] > 0):
                t = h[a] / (h[a] - h[b])
                pts.append(UNIT_TET[a] + t * (UNIT_TET[b] - UNIT_TET[a]))
        pts = np.array(pts)
        c = pts.mean(0)
        ang = np.arctan2((pts - c) @ np.cross(z, x), (pts - c) @ x)
        P = pts[np.argsort(ang)]
        area = 0.5 * abs(sum(np.cross(P[i] - P[0], P[i + 1] - P[0]) @ n for i in range(1, len(P) - 1)))
        assert r["n_polygons"] == 1 and np.isclose(r["area"], area, rtol=1e-12)
        assert s.pair_emitted(0)[0, 2] == len(pts)
        faces = s.pair_faces(0)
        # plane is M here only if its id is lower: plane id 1 > soft id 0, so soft is M and normals point into it
        if len(faces):
            assert np.allclose(faces[0, 3:6], n, atol=1e-14)
    assert seen == set(range(1, 15))


def test_equal_pressure_plane_of_two_linear_fields():
    """Two overlapping unit tets with fields f0 = z and f1 = 1 - z - 0.5 shifted: the contact polygon lies
    on the plane where the fields are equal."""
    s = OracleScene()
    big = np.array([[-2, -2, -2], [4, -2, -2], [-2, 4, -2], [-2, -2, 4.0]])
    a = s.add_raw_soft(big, [[0, 1, 2, 3]], big[:, 2] + 2.0, [1.0, 0, 0, 0, 0])  # f0 = z + 2
    b = s.add_raw_soft(big, [[0, 1, 2, 3]], 2.5 - big[:, 2], [1.0, 0, 0, 0, 0])  # f1 = 2.5 - z
    s.set_pairs([[a, b]])
    s.step(np.zeros((2, 3)), np.stack([I3, I3]), use_bvh=False)
    r = s.pair_result(0)
    assert r["has_surface"] and r["n_polygons"] == 1
    tri_z = 0.25  # z + 2 = 2.5 - z
    assert np.isclose(r["centroid"][2], tri_z, rtol=1e-13)
    # cross-section of the tet x,y,z >= -2, x+y+z <= 0 at z = 0.25: right triangle with legs 3.75
    assert np.isclose(r["area"], 0.5 * 3.75 ** 2, rtol=1e-13)
    faces = s.pair_faces(0)
    assert np.allclose(faces[0, 3:6], [0, 0, 1])  # normal along increasing f0, into geometry 0 (M)
    assert np.isclose(faces[0, 6], r["area"] * 2.25)  # fn0 = area * pressure on the surface
    assert np.isclose(faces[0, 7], r["area"] * 0.5)  # g = 1 / (1/1 + 1/1) = 0.5


# ---- invariances ----------------------------------------------------------------------------------------
def sphere_box_scene(triangle=False):
    s = OracleScene(triangle_representation=triangle)
    box = s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], RIGID)
    sph = s.add_geom(GEOM_SPHERE, [0.08], SOFT)
    s.set_pairs([[sph, box]])
    return s


def test_bvh_and_brute_force_emit_the_same_set():
    s = sphere_box_scene()
    rng = np.random.default_rng(5)
    for _ in range(10):
        R = rot(rng.normal(size=3), rng.uniform(0, np.pi))
        xpos = np.array([[0, 0, 0.1], [rng.uniform(-.05, .05), rng.uniform(-.05, .05), 0.28 - rng.uniform(0.004, 0.03)]])
        xmat = np.stack([I3, R.reshape(-1)])
        s.step(xpos, xmat, use_bvh=True)
        a, ra = set(map(tuple, s.pair_emitted(0).tolist())), s.pair_result(0)
        s.step(xpos, xmat, use_bvh=False)
        b, rb = set(map(tuple, s.pair_emitted(0).tolist())), s.pair_result(0)
        assert a == b and len(a) > 0
        assert ra["n_candidates"] < rb["n_candidates"] == 128 * 48
        assert np.allclose(ra["F"], rb["F"], rtol=1e-12)


def test_rigid_motion_invariance_and_newton_third_law():
    s = sphere_box_scene()
    xpos = np.array([[0, 0, 0.1], [0.01, -0.02, 0.27]])
    R1 = rot([1, 2, 3], 0.7)
    xmat = np.stack([I3, R1.reshape(-1)])
    vel = np.array([[0, 0, 0, 0, 0, 0], [0.1, 0.2, -0.3, 0.01, 0.02, -0.05]])
    s.step(xpos, xmat, vel)
    r0 = s.pair_result(0)
    w = [s.geom_wrench(0), s.geom_wrench(1)]
    assert np.allclose(w[0] + w[1], 0, atol=1e-12)  # equal and opposite wrenches about the world origin
    assert np.allclose(w[0][:3], r0["F"])  # box has the lower id: it is M
    # move the whole scene rigidly (poses AND velocities): force/torque rotate with it
    Q, t = rot([0.3, -1, 0.5], 1.1), np.array([0.4, -0.2, 0.7])
    xpos2 = xpos @ Q.T + t
    xmat2 = np.stack([(Q @ I3.reshape(3, 3)).reshape(-1), (Q @ R1).reshape(-1)])
    vel2 = np.concatenate([vel[:, :3] @ Q.T, vel[:, 3:] @ Q.T], axis=1)
    s.step(xpos2, xmat2, vel2)
    r1 = s.pair_result(0)
    assert r1["n_polygons"] == r0["n_polygons"]
    assert np.allclose(r1["F"], Q @ r0["F"], rtol=1e-9, atol=1e-12)
    assert np.allclose(r1["centroid"], Q @ r0["centroid"] + t, rtol=1e-10)
    assert np.allclose(r1["tau"], Q @ r0["tau"] + np.cross(t, Q @ r0["F"]), rtol=1e-9, atol=1e-12)


def test_pair_order_and_id_order_conventions():
    """(g1,g2) order handed to collision_cb does not matter for soft-rigid; M is the lower config index."""
    s1, s2 = sphere_box_scene(), sphere_box_scene()
    s2.set_pairs([[0, 1]])
    xpos = np.array([[0, 0, 0.1], [0.0, 0.0, 0.27]])
    xmat = np.stack([I3, rot([0, 1, 1], 0.3).reshape(-1)])
    s1.step(xpos, xmat)
    s2.step(xpos, xmat)
    a, b = s1.pair_result(0), s2.pair_result(0)
    assert (a["gM"], a["gN"]) == (0, 1) == (b["gM"], b["gN"])
    assert np.array_equal(a["F"], b["F"]) and a["F"][2] < 0  # sphere pushes the box (M) down
    faces = s1.pair_faces(0)
    assert (faces[:, 5] < 0).all()  # normals point out of the sphere (N) into the box (M)


def test_sphere_on_plane_force_converges_to_closed_form():
    """F = pi E d^2 (1 - 2d/(3R)) ... for p = E (1 - r/R) integrated over the cap at depth d; exact value by
    quadrature of the analytic field, mesh refinement must converge towards it from below."""
    R, E, d = 0.08, 5e4, 0.012
    rho = np.linspace(0, np.sqrt(R * R - (R - d) ** 2), 200001)
    p = E * (1 - np.sqrt(rho ** 2 + (R - d) ** 2) / R)
    exact = np.trapezoid(p * 2 * np.pi * rho, rho)
    assert np.isclose(exact, np.pi * E * d * d * (1 - 2 * d / (3 * R)) , rtol=1e-6)  # closed form, SURVEY §4
    errs = []
    for hint in (0.05, 0.02, 0.01, 0.005):
        s = OracleScene()
        plane = s.add_geom(GEOM_PLANE, [0, 0, 1], [0, 1, 0, 0.3, 0.3])
        sph = s.add_geom(GEOM_SPHERE, [R], [E, 0, hint, 0.3, 0.3])
        s.set_pairs([[plane, sph]])
        s.step(np.array([[0, 0, 0], [0, 0, R - d]]), np.stack([I3, rot([1, 1, 0], 0.4).reshape(-1)]))
        F = s.pair_result(0)["F"]
        errs.append(abs(-F[2] - exact) / exact)  # plane is M (id 0): the sphere pushes it down
        assert abs(F[0]) < 1e-9 and abs(F[1]) < 1e-9
    assert errs[0] > errs[1] > errs[2] > errs[3] and errs[3] < 0.01


def test_force_law_damping_and_friction():
    """plugin.cpp:458-475: fn = max(0, 1 - d vn)(fn0 - 0.001 k vn); regularised Coulomb friction."""
    tri = np.array([[-5, -5, 0.25], [5, -5, 0.25], [0, 8, 0.25]])
    s = OracleScene()
    soft = s.add_raw_soft(UNIT_TET, [[0, 1, 2, 3]], [0, 0, 0, 1.0], [2.0, 3.0, 0, 0.5, 0.5])
    rigid = s.add_raw_rigid(tri, [[0, 1, 2]], [0, 1, 0, 0.5, 0.5])
    s.set_pairs([[soft, rigid]])
    area, p0, g = 0.28125, 0.25, 1.0
    for vz, vx in [(0.0, 0.0), (-0.1, 0.0), (0.5, 0.0), (0.0, 1.0), (-0.2, 5e-5)]:
        vel = np.zeros((2, 6))
        vel[0, 3:] = [vx, 0, vz]  # soft tet (A = M) moves, no rotation
        s.step(np.zeros((2, 3)), np.stack([I3, I3]), vel)
        F = s.pair_result(0)["F"]
        vn = vz  # normal +z into M
        fn = max(0.0, 1 - 3.0 * vn) * (area * p0 - 0.001 * area * g * vn)
        mu = 0.5
        vslip = np.sqrt(vx * vx + (1e-4 * 1e-2) ** 2)
        sreg = vslip / 1e-4
        mu_r = mu * sreg * (2 - sreg) if sreg < 1 else mu
        assert np.allclose(F, [-mu_r * vx / vslip * fn, 0, fn], rtol=1e-12, atol=1e-15)


# ---- flat tactile sensor --------------------------------------------------------------------------------
def myrmex_scene(S=4, window=0, sigma=-1.0):
    s = OracleScene(triangle_representation=True)
    box = s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1, 0.05, 0.3, 0.3])
    foam = s.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5, 0, 0.3, 0.3])
    s.set_pairs([[box, foam]])
    s.add_flat_sensor(foam, [0.2, 0.2, 0.02], 0.025, S, window, sigma)
    return s


def test_flat_sensor_dims_and_flat_press_image():
    s = myrmex_scene()
    assert s.sensor_dims(0) == (16, 16)
    depth = 0.002
    xpos = np.array([[0, 0, 0.053 - depth + 0.1], [0, 0, 0.033]])
    s.step(xpos, np.stack([I3, I3]))
    img = s.sensor_image(0).reshape(16, 16)
    img_brute = s.sensor_image(0, use_bvh=False).reshape(16, 16)
    # box 0.2 x 0.2 centred on a 0.4 x 0.4 pad: exactly the 8x8 central taxels feel p = E * depth / 0.02
    inner = img[4:12, 4:12]
    assert np.allclose(inner, 5e4 * depth / 0.02, rtol=1e-5)
    assert np.count_nonzero(img) == 64
    assert np.allclose(img, img_brute, rtol=1e-6)
    # the image integrates to the normal force: sum(p * taxel area) = F
    F = s.pair_result(0)["F"]
    assert np.isclose(img.sum() * 0.025 ** 2, abs(F[2]), rtol=1e-5)


def test_flat_sensor_zero_without_contact_and_window_weights():
    s = myrmex_scene(S=8, window=1, sigma=0.1)
    s.step(np.array([[0, 0, 0.5], [0, 0, 0.033]]), np.stack([I3, I3]))
    assert not s.sensor_image(0).any()
    s.step(np.array([[0, 0, 0.151], [0, 0, 0.033]]), np.stack([I3, I3]))
    g = s.sensor_image(0).reshape(16, 16)
    s0 = myrmex_scene(S=8)
    s0.step(np.array([[0, 0, 0.151], [0, 0, 0.033]]), np.stack([I3, I3]))
    plain = s0.sensor_image(0).reshape(16, 16)
    # gaussian window: every weight <= 1, so each taxel is attenuated by the same factor (uniform pressure)
    ratio = g[4:12, 4:12] / plain[4:12, 4:12]
    assert (ratio < 1).all() and np.allclose(ratio, ratio[0, 0], rtol=1e-5)


def test_curved_sensor_assignment_weights_and_flat_press():
    """CurvedSensor (curved_sensor.cpp): samples within include_margin of a taxel (and within 45 degrees of its
    normal) are assigned with weight (margin - distance)^2; under a flat press of depth d every ray that starts on
    the foam's top face inside the contact patch reads p = E d / zs."""
    s = OracleScene(triangle_representation=True)
    box = s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1, 0.05, 0.3, 0.3])
    foam = s.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5, 0, 0.3, 0.3])
    s.set_pairs([[box, foam]])
    margin = 0.005
    # (positions off the mesh's symmetry axes: a ray that starts exactly above a triangle edge can slip between the two
    # float32 Moeller-Trumbore tests, in the reference as well)
    taxels = np.array([[0.0113, 0.0071, 0.02], [0.0513, 0.0037, 0.02], [0.19, 0.19, 0.02]])  # third: outside the patch
    normals = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]])
    offs = np.array([[0, 0], [0.001, 0], [0, 0.002], [-0.003, 0], [0.0, -0.0045], [0.004, 0.004]])  # last: 5.66 mm away
    offs = offs + 1e-6 * np.array([[0.3, 0.7]])  # every sample a hair off the taxel: distances stay as listed to 1e-3
    samples = np.concatenate([np.c_[t[0] + offs[:, 0], t[1] + offs[:, 1], np.full(len(offs), 0.02)] for t in taxels])
    snorm = np.tile([0, 0, 1.0], (len(samples), 1))
    snorm[1] = [0, 1, 0]  # 90 degrees off the taxel normal: rejected by the 45 degree test
    cs = s.add_curved_sensor(foam, taxels, normals, samples, snorm, margin)
    n_close, n_assign = s.curved_info(cs)
    assert (n_close, n_assign) == (3 * 5 - 1, 3 * 5 - 1)
    depth = 0.002
    s.step(np.array([[0, 0, 0.053 - depth + 0.1], [0, 0, 0.033]]), np.stack([I3, I3]))
    v = s.curved_values(cs)
    p = 5e4 * depth / 0.02
    d = np.linalg.norm(offs[:5], axis=1)
    w = (margin - d) ** 2
    assert np.isclose(v[0], p * (w.sum() - w[1]), rtol=1e-5)   # sample 1 of taxel 0 was rejected
    assert np.isclose(v[1], p * w.sum(), rtol=1e-5)
    assert v[2] == 0                                            # no contact under that taxel
    assert np.allclose(v, s.curved_values(cs, use_bvh=False), rtol=1e-6)
    # without taxel normals the 45 degree test is off
    s2 = OracleScene(triangle_representation=True)
    b2 = s2.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1, 0.05, 0.3, 0.3])
    f2 = s2.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5, 0, 0.3, 0.3])
    s2.set_pairs([[b2, f2]])
    c2 = s2.add_curved_sensor(f2, taxels, None, samples, snorm, margin)
    assert s2.curved_info(c2) == (15, 15)


def _taxel_reference(tris, taxel_w, margin, res, method, visualize, previous):
    """Independent numpy restatement of TaxelSensor::internal_update (default sampling) for one taxel."""
    pts, prs = [], []
    for t in tris:
        v0, v1, v2, e = t[0:3], t[3:6], t[6:9], t[9:12]
        st0 = int(np.linalg.norm(v1 - v0) / res) + 1
        st1 = int(np.linalg.norm(v2 - v0) / res) + 1
        st, st2 = max(st0, st1), st1
        a = 0.0
        while a <= 1:
            b = 0.0
            while b <= 1:
                bary = np.array([a, (1 - a) * (1 - b), (1 - a) * b])
                pts.append(bary[0] * v0 + bary[1] * v1 + bary[2] * v2)
                prs.append(bary @ e)
                b += 1. / st2
            a += 1. / st
    if not pts:
        return 0.0
    pts, prs = np.array(pts), np.array(prs)
    d2 = ((pts - taxel_w) ** 2).sum(1)
    if method == "closest":
        j = d2.argmin()
        if d2[j] < margin ** 2:
            return prs[j] if (visualize and abs(prs[j]) > 1e-6) else 0.0
        return previous
    sel = d2 < margin ** 2
    if not sel.any():
        return previous
    return res * (((margin - np.sqrt(d2[sel])) ** 2) * np.abs(prs[sel])).sum()


def test_taxel_sensor_methods_and_their_quirks():
    s = OracleScene(triangle_representation=True)
    box = s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1, 0.05, 0.3, 0.3])
    foam = s.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5, 0, 0.3, 0.3])
    s.set_pairs([[box, foam]])
    taxels = np.array([[0.013, 0.007, 0.02], [0.09, -0.05, 0.02], [0.19, 0.19, 0.02]])  # the last one is out of reach
    margin, res = 0.0126, 0.01
    ids = {m: s.add_taxel_sensor(foam, taxels, margin, res, m, vis)
           for m, vis in (("weighted", False), ("mean", False), ("squared", False), ("closest", False))}
    ids["closest_vis"] = s.add_taxel_sensor(foam, taxels, margin, res, "closest", True)
    R = np.array([[np.cos(0.3), -np.sin(0.3), 0], [np.sin(0.3), np.cos(0.3), 0], [0, 0, 1]])
    s.step(np.array([[0.01, 0.02, 0.053 - 0.002 + 0.1], [0, 0, 0.033]]), np.stack([R.reshape(-1), I3]))
    tris = s.pair_triangles(0)
    assert len(tris) > 10
    prev = np.array([7.0, 8.0, 9.0], dtype=np.float32)
    world = taxels + [0, 0, 0.033]
    out = {k: s.taxel_values(i, previous=prev) for k, i in ids.items()}
    # the missing breaks: weighted and mean end with the squared result
    assert np.array_equal(out["weighted"], out["squared"]) and np.array_equal(out["mean"], out["squared"])
    for k, (m, vis) in {"squared": ("squared", False), "closest": ("closest", False), "closest_vis": ("closest", True)}.items():
        ref = [_taxel_reference(tris, world[i], margin, res, m, vis, float(prev[i])) for i in range(3)]
        assert np.allclose(out[k], ref, rtol=1e-6), (k, out[k], ref)
    assert out["squared"][2] == 9.0 and out["closest"][2] == 9.0      # no sample in range: the old value stays
    assert out["closest"][0] == 0.0 and out["closest_vis"][0] > 100   # closest keeps the pressure only with visualize
    # no sample at all: zeros
    s.step(np.array([[0, 0, 0.5], [0, 0, 0.033]]), np.stack([I3, I3]))
    assert not s.taxel_values(ids["squared"], previous=prev).any()


def _minstd_canonical(state):
    """std::generate_canonical<double, 53>(std::minstd_rand0) as libstdc++ evaluates it (two 31-bit draws per double);
    independent of the oracle, which calls the C++ standard library itself."""
    r = 2147483646.0
    state = (16807 * state) % 2147483647
    s = float(state - 1)
    state = (16807 * state) % 2147483647
    s += float(state - 1) * r
    u = s / (r * r)
    return (np.nextafter(1.0, 0.0) if u >= 1.0 else u), state


def _area_importance_reference(tris, emitted, taxel_w, margin, res, previous):
    """taxel_sensor.cpp:211-254 + the squared method, in numpy: triangles in the canonical order (polygons by
    (elemM, elemN), fan triangles in fan order), strata of res * total_area, one minstd_rand0 stream."""
    order, first = [], 0
    firsts = []
    for em in emitted:
        firsts.append(first)
        first += int(em[2])
    for q in sorted(range(len(emitted)), key=lambda q: (int(emitted[q][0]), int(emitted[q][1]))):
        order += list(range(firsts[q], firsts[q] + int(emitted[q][2])))
    area_of = lambda t: 0.5 * np.linalg.norm(np.cross(t[3:6] - t[0:3], t[6:9] - t[0:3]))
    total = 0.0
    for i in order:
        total += area_of(tris[i])
    step, acc, at, state = res * total, 0.0, 0.0, 1
    pts, prs = [], []
    for i in order:
        t = tris[i]
        at += area_of(t)
        while acc < at:
            acc += step
            u0, state = _minstd_canonical(state)
            u1, state = _minstd_canonical(state)
            a = 1.0 - np.sqrt(u0)
            b = (1.0 - a) * u1
            bary = np.array([a, (1 - a) * (1 - b), (1 - a) * b])
            pts.append(bary[0] * t[0:3] + bary[1] * t[3:6] + bary[2] * t[6:9])
            prs.append(bary @ t[9:12])
    pts, prs = np.array(pts), np.array(prs)
    d2 = ((pts - taxel_w) ** 2).sum(1)
    sel = d2 < margin ** 2
    if not sel.any():
        return previous, len(pts)
    return res * (((margin - np.sqrt(d2[sel])) ** 2) * np.abs(prs[sel])).sum(), len(pts)


def test_taxel_sensor_area_importance_sampling():
    """sample_method "area_importance" (the one the reference's fingertip.yaml uses): oracle (std::default_random_engine)
    against an independent numpy restatement with its own minstd_rand0 / generate_canonical."""
    s = OracleScene(triangle_representation=True)
    box = s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1, 0.05, 0.3, 0.3])
    foam = s.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5, 0, 0.3, 0.3])
    s.set_pairs([[box, foam]])
    taxels = np.array([[0.013, 0.007, 0.02], [0.09, -0.05, 0.02], [0.19, 0.19, 0.02]])
    margin, res = 0.0126, 0.002
    sid = s.add_taxel_sensor(foam, taxels, margin, res, "squared", False, "area_importance")
    R = np.array([[np.cos(0.3), -np.sin(0.3), 0], [np.sin(0.3), np.cos(0.3), 0], [0, 0, 1]])
    s.step(np.array([[0.01, 0.02, 0.053 - 0.002 + 0.1], [0, 0, 0.033]]), np.stack([R.reshape(-1), I3]))
    tris, emitted = s.pair_triangles(0), s.pair_emitted(0)
    assert len(tris) == emitted[:, 2].sum() > 10
    prev = np.array([7.0, 8.0, 9.0], dtype=np.float32)
    out = s.taxel_values(sid, previous=prev)
    world = taxels + [0, 0, 0.033]
    ref = [_area_importance_reference(tris, emitted, world[i], margin, res, float(prev[i])) for i in range(3)]
    assert 500 <= ref[0][1] <= 502  # one sample per stratum of res * total_area
    assert np.allclose(out, [r[0] for r in ref], rtol=1e-6), (out, ref)
    assert out[2] == 9.0 and out[0] > 0
    # the generator restarts every update: the same contact gives the same message
    assert np.array_equal(out, s.taxel_values(sid, previous=prev))


@pytest.mark.parametrize("triangle", [False, True])
def test_face_vertices_export_is_consistent_with_the_point_collisions(triangle):
    """orc_pair_face_vertices (the faces visualizeMeshElement outlines, plugin.cpp:525-555): every face's area-weighted
    centroid is its PointCollision's p, its right-hand normal is n, and area x pressure at the centroid is fn0."""
    s = OracleScene(triangle_representation=triangle)
    box = s.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1, 0.1, 0.3, 0.3])
    sph = s.add_geom(GEOM_SPHERE, [0.08], [5e4, 5, 0.05, 0.3, 0.3])
    s.set_pairs([[sph, box]])
    R = np.array([[np.cos(0.4), 0, np.sin(0.4)], [0, 1, 0], [-np.sin(0.4), 0, np.cos(0.4)]])
    s.step(np.array([[0, 0, 0.1], [0.01, -0.02, 0.2 + 0.08 - 0.012]]), np.stack([I3, R.reshape(-1)]))
    nv, fv = s.pair_face_vertices(0)
    pcs = s.pair_faces(0)
    assert len(nv) == len(pcs) > 10 and (nv == 3).all() == triangle
    total = 0.0
    for k in range(len(nv)):
        v = fv[k, :nv[k]]
        a2 = [np.cross(v[i] - v[0], v[i + 1] - v[0]) for i in range(1, nv[k] - 1)]
        nrm = np.sum(a2, axis=0)
        area = 0.5 * np.linalg.norm(nrm)
        cen = sum(np.linalg.norm(a) * (v[0] + v[i + 1] + v[i + 2]) / 3 for i, a in enumerate(a2)) / (2 * area)
        assert np.allclose(cen, pcs[k, 0:3], atol=1e-12)
        assert np.allclose(nrm / (2 * area), pcs[k, 3:6], atol=1e-9)
        total += area
    assert abs(total - s.pair_result(0)["area"]) < 1e-12


def soft_box_stack_closed_form(E, a, b, c, d):
    """Two soft medial-axis boxes of equal thickness 2c stacked with overlap d, the upper one wider than the lower one's
    footprint 2a x 2b by more than d/2.  Fields: p0 = E/c min(a-|x|, b-|y|, dist to the top face), p1 = E/c (dist to the
    upper box's bottom face) over the footprint.  The equal-pressure surface is the mid-plane in the interior and a 45-degree
    roof within t = d/2 of the lower box's sides; its pressure over the footprint is E/c min(t, a-|x|, b-|y|), so by the
    layer-cake formula  F_z = E/c int_0^t (2a-2s)(2b-2s) ds = E/c (4ab t - 2(a+b) t^2 + 4/3 t^3)  (n_z dA = dx dy),
    and the surface area is the inner rectangle plus sqrt(2) times the frame."""
    t = d / 2
    force = E / c * (4 * a * b * t - 2 * (a + b) * t * t + 4 / 3 * t ** 3)
    inner = (2 * a - 2 * t) * (2 * b - 2 * t)
    return force, inner + np.sqrt(2.0) * (4 * a * b - inner)


@pytest.mark.parametrize("triangle", [False, True])
@pytest.mark.parametrize("d", [0.004, 0.01, 0.016])
def test_soft_box_stack_force_and_area_closed_form(d, triangle):
    """Soft-soft path end to end (medial-axis box meshes, BVH, equal-pressure planes incl. the oblique pieces where one field
    is governed by a side face, both gradient culls, clip, quadrature, force law at rest): the linear fields are exact on the
    meshes, so the closed form must be met to rounding, in any rigid placement of the pair."""
    E, a, b, c, A, B = 5e4, 0.06, 0.04, 0.02, 0.12, 0.10
    force, area = soft_box_stack_closed_form(E, a, b, c, d)
    for R, p in ((np.eye(3), np.zeros(3)), (rot([1, 2, 3], 0.7), np.array([0.3, -0.2, 0.5]))):
        s = OracleScene(triangle_representation=triangle)  # kTriangle: centroid fans, pressure (ea + eb + ec) / 3: exact too
        g0 = s.add_geom(GEOM_BOX, [a, b, c], [E, 0, 0, 0.3, 0.3])
        g1 = s.add_geom(GEOM_BOX, [A, B, c], [E, 0, 0, 0.3, 0.3])
        s.set_pairs([[g0, g1]])
        offs = np.array([0.003, -0.002, 2 * c - d])
        s.step(np.array([p, p + R @ offs]), np.stack([R.reshape(-1), R.reshape(-1)]))
        r = s.pair_result(0)
        assert r["has_surface"]
        F_local = R.T @ np.asarray(r["F"])
        assert np.isclose(abs(F_local[2]), force, rtol=1e-12)
        assert abs(F_local[0]) < 1e-12 * force and abs(F_local[1]) < 1e-12 * force
        assert np.isclose(r["area"], area, rtol=1e-12)
        # the surface is symmetric about the lower box's axis: so is its area centroid.  (The torque about that axis is NOT
        # zero to rounding: one quadrature point per polygon integrates the linear pressure exactly, its moment only to
        # second order, and the medial-axis tetrahedralisation is not symmetric: a lever arm below 1 % of the half size here.)
        cen_local = R.T @ (np.asarray(r["centroid"]) - p)
        assert abs(cen_local[0]) < 1e-12 and abs(cen_local[1]) < 1e-12 and c - d < cen_local[2] < c
        tau_about_box0 = np.asarray(r["tau"]) - np.cross(p, np.asarray(r["F"]))
        assert np.linalg.norm(tau_about_box0) < 1e-2 * force * a


@pytest.mark.parametrize("d", [0.004, 0.01, 0.016])
def test_crossed_soft_boxes_force_closed_form(d):
    """As above with the upper box NARROWER than the lower one in x and wider in y: within t = d/2 of the upper box's sides
    ITS field is governed by a side face (roof pieces whose normal leans the other way), in the corners both are (a vertical
    piece of the equal-pressure surface, which carries no F_z).  Over the common footprint 2a' x 2b the surface pressure is
    E/c min(t, a'-|x|, b-|y|) again, so F_z = E/c (4a'b t - 2(a'+b) t^2 + 4/3 t^3)."""
    E, c = 5e4, 0.02
    lo, up = (0.08, 0.04, c), (0.03, 0.10, c)
    force, _ = soft_box_stack_closed_form(E, up[0], lo[1], c, d)
    for R, p in ((np.eye(3), np.zeros(3)), (rot([1, 2, 3], 0.7), np.array([0.3, -0.2, 0.5]))):
        s = OracleScene()
        g0 = s.add_geom(GEOM_BOX, list(lo), [E, 0, 0, 0.3, 0.3])
        g1 = s.add_geom(GEOM_BOX, list(up), [E, 0, 0, 0.3, 0.3])
        s.set_pairs([[g0, g1]])
        s.step(np.array([p, p + R @ np.array([0.004, -0.003, 2 * c - d])]), np.stack([R.reshape(-1), R.reshape(-1)]))
        F_local = R.T @ np.asarray(s.pair_result(0)["F"])
        assert np.isclose(abs(F_local[2]), force, rtol=1e-12)
        assert abs(F_local[0]) < 1e-12 * force and abs(F_local[1]) < 1e-12 * force


@pytest.mark.parametrize("d", [0.003, 0.008, 0.015])
def test_faces_without_pressure_gradient_along_their_normal_carry_no_force(d):
    """plugin.cpp:366-373 (reference-owned): a face with gM < 1e-14 or gN < 1e-14 is skipped.  A soft medial-axis box pressed d
    into a rigid plane, axis-aligned: the cut is the rectangle 2a x 2b (the surface's area), but within d of the sides the field
    is governed by a side face, its gradient is horizontal with an exactly zero z component (the box's vertex coordinates and
    pressures make it so), the plane's normal is exactly (0, 0, 1), so the box's gradient along the normal (gN: the box is N) is
    exactly 0 and those faces carry no force: what is left is the inner rectangle under the uniform pressure E d / h:
    F = E d / h (2a - 2d)(2b - 2d).  (Pressed by a rigid BOX instead, the face normals carry 1e-16 of rounding, one frame
    face gets gN ~ 1e-10 > 1e-14 and its whole pressure counts: the rule is a knife edge there, in the reference as here.)"""
    E, a, b, c = 5e4, 0.06, 0.04, 0.03
    s = OracleScene()
    plane = s.add_geom(GEOM_PLANE, [0, 0, 1], [0, 1, 0, 0.3, 0.3])
    box = s.add_geom(GEOM_BOX, [a, b, c], [E, 0, 0, 0.3, 0.3])
    s.set_pairs([[plane, box]])
    s.step(np.array([[0, 0, 0], [0.01, 0.02, c - d]]), np.stack([I3, I3]))
    r = s.pair_result(0)
    assert np.isclose(r["area"], 4 * a * b, rtol=1e-13)
    assert np.isclose(abs(r["F"][2]), E * d / min(a, b, c) * (2 * a - 2 * d) * (2 * b - 2 * d), rtol=1e-12)
    assert abs(r["F"][0]) < 1e-12 and abs(r["F"][1]) < 1e-12
    assert 0 < r["n_points"] < r["n_faces"]
