"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs.  Bars (BASELINE.json north star): emitted candidate set + polygon vertex counts bit-exact,
per-pair force/torque within 1e-8 relative (fp64 geometry mode), taxel images within 1e-6 relative."""
import os

import numpy as np
import pytest

from mujoco_contact_surfaces_b200 import scenes
from parity_utils import (TAXEL_RTOL, compare_env, compare_images, compare_wrench, make_engine, make_oracle, oracle_env)

pytestmark = pytest.mark.gpu


def _run_scene(scene, n_envs, seed, hcs_lib, with_sensors=False, check_images=True, **engine_kw):
    eng = make_engine(scene, n_envs, **engine_kw)
    orc = make_oracle(scene)
    xpos, xmat, vel = scene.poses(n_envs, seed)
    eng.step(xpos, xmat, vel, with_sensors=with_sensors)
    res = eng.pair_results()
    wrench = eng.geom_wrenches()
    imgs = [eng.sensor_image(s) for s in range(len(scene.sensors))] if with_sensors else []
    worst, n_poly = 0.0, 0
    for e in range(n_envs):
        ref_pairs, ref_imgs = oracle_env(orc, scene, xpos[e], xmat[e], vel[e], sensors=with_sensors)
        emitted = [eng.emitted(e, p) for p in range(len(scene.pairs))]
        worst = max(worst, compare_env(res[e], emitted, ref_pairs))
        n_poly += sum(r["n_polygons"] for r in ref_pairs)
        for g in range(scene.n_geoms):
            compare_wrench(wrench[e, g], orc.geom_wrench(g))
        if with_sensors and check_images:
            for s, ref_img in enumerate(ref_imgs):
                err, nbad = compare_images(imgs[s][e], ref_img)
                assert nbad == 0, "taxel image: %d taxels beyond %.0e (max rel err %.3e)" % (nbad, TAXEL_RTOL, err)
    assert n_poly > 0, "scene produced no contact at all: the test would be vacuous"
    eng.close()
    return worst


def test_meshes_match_oracle_bit_exact(hcs_lib):
    """Stage (1): device-resident meshes, pressures, gradients and normals equal the oracle's bit for bit."""
    for name, fn in scenes.SCENES.items():
        scene = fn()
        eng, orc = make_engine(scene, 1), make_oracle(scene)
        for g in range(scene.n_geoms):
            a, b = eng.geom_mesh(g), orc.geom_mesh(g)
            assert a["kind"] == b["kind"], (name, g)
            if a["kind"] == 2:
                continue
            for key in b:
                if key == "kind":
                    continue
                assert np.array_equal(a[key], b[key]), "%s geom %d: %s differs" % (name, g, key)
        eng.close()


def test_gpu_lbvh_build_equals_host_builder(hcs_lib, monkeypatch):
    """K2: Morton sort + Karras tree + atomic refit on the GPU give the same 64-byte node records as the host
    cross-check builder, and every tet is reachable exactly once with a box that contains it."""
    scene = scenes.mixed_shapes()
    scene.geoms.append(scenes.Geom("ell_fine", scenes.GEOM_ELLIPSOID, [0.05, 0.04, 0.03], [5e4, 5.0, 0.004, 0.3, 0.3]))
    gpu = make_engine(scene, 1)
    monkeypatch.setenv("HCS_LBVH_HOST", "1")
    host = make_engine(scene, 1)
    monkeypatch.delenv("HCS_LBVH_HOST")
    checked = 0
    for g in range(scene.n_geoms):
        m = gpu.geom_mesh(g)
        if m["kind"] != 1:
            continue
        a, b = gpu.lbvh(g), host.lbvh(g)
        assert a.tobytes() == b.tobytes(), "LBVH of geom %d differs between GPU and host builders" % g
        ntet = len(m["elems"])
        leaves = np.concatenate([a["left"][a["left"] < 0], a["right"][a["right"] < 0]])
        assert sorted((~leaves).tolist()) == list(range(ntet))
        for side, lo, hi in (("left", "llo", "lhi"), ("right", "rlo", "rhi")):
            sel = a[side] < 0
            tv = m["verts"][m["elems"][~a[side][sel]]]
            assert (tv.min(1) >= a[lo][sel] - 1e-12).all() and (tv.max(1) <= a[hi][sel] + 1e-12).all()
        checked += 1
    assert checked >= 5
    gpu.close(), host.close()


def test_c1_sphere_on_box_random_orientation(hcs_lib):
    worst = _run_scene(scenes.sphere_on_box(), 128, seed=1234, hcs_lib=hcs_lib)
    assert worst < 1e-8


def test_c1_identity_orientation_degenerate_prone(hcs_lib):
    """Axis-aligned resting poses put triangle vertices exactly on tet faces; reported separately."""
    _run_scene(scenes.sphere_on_box(identity_orientation=True), 32, seed=99, hcs_lib=hcs_lib)


def test_c3_soft_soft_polygon(hcs_lib):
    _run_scene(scenes.soft_soft(hint=0.01), 48, seed=3, hcs_lib=hcs_lib)


def test_c3_soft_soft_triangle(hcs_lib):
    _run_scene(scenes.soft_soft(hint=0.02, triangle=True), 16, seed=4, hcs_lib=hcs_lib)


def test_c4_objects_on_plane(hcs_lib):
    _run_scene(scenes.objects_on_plane(), 64, seed=4096, hcs_lib=hcs_lib)


def test_c4_objects_on_plane_triangle(hcs_lib):
    _run_scene(scenes.objects_on_plane(triangle=True), 16, seed=4097, hcs_lib=hcs_lib)


@pytest.mark.parametrize("presser,S", [("box", 4), ("box", 20), ("plate", 8), ("spot", 8), ("soft_tip", 8), ("soft_tip", 20)])
def test_c2_myrmex_taxel_image(hcs_lib, presser, S):
    # soft_tip = C2b: a SOFT convex-mesh presser (soft-soft query feeding the flat sensor)
    _run_scene(scenes.myrmex(presser, sampling_resolution=S), 8, seed=7, hcs_lib=hcs_lib, with_sensors=True)


@pytest.mark.parametrize("presser,resolution,S", [("box", 0.025, 32), ("spot", 0.025, 32), ("plate", 0.0025, 4),
                                                  ("spot", 0.0025, 8)])
def test_c2_benchmark_flat_grid_corners(hcs_lib, presser, resolution, S):
    """The corners of the reference's own benchmark grid (benchmark_flat.cpp:282-373): 32 x 32 rays per taxel, and the
    160 x 160 taxel image of resolution 0.0025."""
    _run_scene(scenes.myrmex(presser, sampling_resolution=S, resolution=resolution), 2, seed=7, hcs_lib=hcs_lib,
               with_sensors=True)


@pytest.mark.parametrize("window,sigma", [(1, 0.1), (2, 0.3), (3, -1.0)])
def test_c2_myrmex_windows(hcs_lib, window, sigma):
    _run_scene(scenes.myrmex("box", sampling_resolution=8, window=window, sigma=sigma), 4, seed=8, hcs_lib=hcs_lib,
               with_sensors=True)


@pytest.mark.parametrize("obj", ["box", "spot"])
def test_c5_grasp_five_pads_with_tactile_arrays(hcs_lib, obj):
    """C5 at reduced pad resolution (level 4: 2048 tets per pad): five soft pads on one rigid object, five pairs,
    one 16 x 16 flat sensor per pad whose frame is the pad's own (rays start beyond the pad centre)."""
    _run_scene(scenes.grasp(obj, pad_hint=0.002), 6, seed=55, hcs_lib=hcs_lib, with_sensors=True)


def test_c5_grasp_full_resolution_pads(hcs_lib):
    """C5 at full pad resolution (level 7: 131072 tets, LBVH built on the GPU, ~10^4 polygons per pad)."""
    scene = scenes.grasp("box", n_pads=2)
    eng = make_engine(scene, 1)
    assert len(eng.geom_mesh(1)["elems"]) == 131072
    eng.close()
    _run_scene(scene, 2, seed=56, hcs_lib=hcs_lib, with_sensors=True)


@pytest.mark.parametrize("with_normals", [True, False])
def test_curved_fingertip_sensor(hcs_lib, with_normals):
    """CurvedSensor on the soft ubi_tip fingertip (reference taxel layout): sample-to-taxel assignment and the
    per-taxel weighted ray pressures against the oracle's restatement of curved_sensor.cpp."""
    scene = scenes.fingertip(with_normals=with_normals)
    n_envs = 12
    eng, orc = make_engine(scene, n_envs), make_oracle(scene)
    n_tax, n_rays, n_assign = eng.curved_info(0)
    assert (n_rays, n_assign) == orc.curved_info(0) and n_tax == 12 and n_rays > 500
    xpos, xmat, vel = scene.poses(n_envs, seed=9)
    eng.step(xpos, xmat, vel, with_sensors=True)
    vals = eng.curved_values(0)
    res = eng.pair_results()
    touched = 0
    for e in range(n_envs):
        ref_pairs, _ = oracle_env(orc, scene, xpos[e], xmat[e], vel[e], sensors=False)
        compare_env(res[e], [eng.emitted(e, 0)], ref_pairs)
        ref = orc.curved_values(0)
        err, nbad = compare_images(vals[e], ref)
        assert nbad == 0, "curved sensor: %d taxels beyond %.0e (max rel err %.3e)" % (nbad, TAXEL_RTOL, err)
        touched += int((ref > 0).sum())
    assert touched >= n_envs, "the taxels barely respond: the test would be vacuous"
    eng.close()


@pytest.mark.parametrize("presser,method,visualize,sample_method",
                         [("box", "weighted", False, "default"), ("spot", "squared", False, "default"),
                          ("box", "closest", True, "default"), ("box", "closest", False, "default"),
                          ("box", "squared", False, "area_importance"), ("spot", "squared", False, "area_importance"),
                          ("soft_tip", "closest", True, "area_importance")])
def test_taxel_sensor_on_the_myrmex_foam(hcs_lib, presser, method, visualize, sample_method):
    """TaxelSensor (flat_taxel_sensor.yaml lattice): two consecutive updates with different poses, so that taxels
    that lose their samples keep the previous value like the reference's message buffer does.  area_importance: the
    stratified random sampling of taxel_sensor.cpp:211-254 (soft_tip: a soft-soft surface, M/N swapped windings)."""
    scene = scenes.myrmex_taxels(presser, method, visualize, sample_method,
                                 0.002 if sample_method == "area_importance" else 0.01)
    n_envs = 6
    eng, orc = make_engine(scene, n_envs), make_oracle(scene)
    prev = np.zeros((n_envs, 256), dtype=np.float32)
    moved = 0
    for seed in (31, 32):
        xpos, xmat, vel = scene.poses(n_envs, seed=seed)
        eng.step(xpos, xmat, vel, with_sensors=True)
        vals = eng.taxel_values(0)
        for e in range(n_envs):
            oracle_env(orc, scene, xpos[e], xmat[e], vel[e], sensors=False)
            ref = orc.taxel_values(0, previous=prev[e])
            err, nbad = compare_images(vals[e], ref)
            assert nbad == 0, "taxel sensor: %d taxels beyond %.0e (max rel err %.3e)" % (nbad, TAXEL_RTOL, err)
            moved += int((ref != prev[e]).sum())
            prev[e] = ref
    assert moved > 0 or (method == "closest" and not visualize)  # closest without visualize writes zeros (Q12)
    eng.close()


@pytest.mark.parametrize("sample_method", ["default", "area_importance"])
def test_taxel_sensor_on_the_fingertip(hcs_lib, sample_method):
    """SENS/config/fingertip.yaml on the soft ubi_tip mesh: its taxels, method squared, sample_resolution 0.001 and (second
    case) its sample_method area_importance, next to the curved sensor."""
    scene = scenes.fingertip()
    scene.taxel_sensors = [dict(geom=1, taxel_pos=scenes._TIP_TAXELS, include_margin=0.006, sample_resolution=0.001,
                                method="squared", sample_method=sample_method)]
    n_envs = 8
    eng, orc = make_engine(scene, n_envs), make_oracle(scene)
    xpos, xmat, vel = scene.poses(n_envs, seed=10)
    eng.step(xpos, xmat, vel, with_sensors=True)
    vals, cvals = eng.taxel_values(0), eng.curved_values(0)
    hit = 0
    for e in range(n_envs):
        oracle_env(orc, scene, xpos[e], xmat[e], vel[e], sensors=False)
        ref = orc.taxel_values(0)
        err, nbad = compare_images(vals[e], ref)
        assert nbad == 0, "taxel sensor: %d taxels beyond %.0e (max rel err %.3e)" % (nbad, TAXEL_RTOL, err)
        assert compare_images(cvals[e], orc.curved_values(0))[1] == 0
        hit += int((ref > 0).sum())
    assert hit >= n_envs
    eng.close()


@pytest.mark.parametrize("triangle", [False, True])
def test_mixed_shapes_every_mesh_family(hcs_lib, triangle):
    """Soft MA cylinders (segment and disc regimes), soft grid box, soft MA cube, rigid box / sphere / ellipsoid /
    cylinder surfaces, a soft-soft cylinder pair and one pair without contact."""
    _run_scene(scenes.mixed_shapes(triangle=triangle), 24, seed=21, hcs_lib=hcs_lib)


def test_geom_update_rebuilds_mesh_and_field(hcs_lib):
    """onGeomChanged (plugin.cpp:828-975): after a size change the engine matches an oracle built at the new size."""
    scene = scenes.sphere_on_box()
    eng = make_engine(scene, 4)
    xpos, xmat, vel = scene.poses(4, seed=77)
    eng.step(xpos, xmat, vel)
    before = eng.pair_results()["F"].copy()
    eng.update_geom(1, [0.085, 0, 0])  # sphere0 grows by 5 mm
    eng.step(xpos, xmat, vel)
    res = eng.pair_results()
    bigger = scenes.sphere_on_box()
    bigger.geoms[1].size[0] = 0.085
    orc = make_oracle(bigger)
    for e in range(4):
        ref_pairs, _ = oracle_env(orc, bigger, xpos[e], xmat[e], vel[e], sensors=False)
        compare_env(res[e], [eng.emitted(e, 0)], ref_pairs)
    assert not np.allclose(before, res["F"])
    eng.close()


def test_batch_is_env_independent(hcs_lib):
    """A shard of envs gives bit-identical results to the same envs inside a larger batch (multi-GPU
    sharding is by env index with no exchange)."""
    scene = scenes.sphere_on_box()
    xpos, xmat, vel = scene.poses(64, seed=5)
    full = make_engine(scene, 64)
    full.step(xpos, xmat, vel)
    r_full = full.pair_results()
    half = make_engine(scene, 32)
    half.step(xpos[32:], xmat[32:], vel[32:])
    r_half = half.pair_results()
    assert r_full[32:].tobytes() == r_half.tobytes()


def test_candidate_pool_overflow_is_reported_not_undefined(hcs_lib, monkeypatch):
    """Capacity errors come back through the C ABI (SURVEY.md 8b "Errors"): a pool that cannot hold the step's
    candidates makes hcs_step fail with HCS_E_CAPACITY and a message naming the knob; a larger pool recovers."""
    from mujoco_contact_surfaces_b200.engine import HcsError
    scene = scenes.sphere_on_box()
    xpos, xmat, vel = scene.poses(64, seed=3)
    monkeypatch.setenv("HCS_MAX_TOTAL_CANDIDATES", "1024")  # 64 envs x ~47 candidates do not fit
    eng = make_engine(scene, 64)
    with pytest.raises(HcsError) as err:
        eng.step(xpos, xmat, vel)
    assert "candidate pool overflow" in str(err.value)
    eng.close()
    monkeypatch.delenv("HCS_MAX_TOTAL_CANDIDATES")
    eng = make_engine(scene, 64)
    eng.step(xpos, xmat, vel)
    assert eng.pair_results()["n_polygons"].sum() > 0
    eng.close()


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: the oracle cannot walk 4096 environments within the suite's budget, so the whole
# batch is checked through properties that do not depend on its size, and a seeded sample of it against the oracle.
# ---------------------------------------------------------------------------------------------------------
def _random_rigid_motion(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    Q = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return Q, rng.uniform(-0.3, 0.3, size=3)


@pytest.mark.parametrize("factory,n_envs,seed", [(scenes.sphere_on_box, 4096, 1234), (scenes.objects_on_plane, 4096, 4096)])
def test_full_size_batch_properties(hcs_lib, factory, n_envs, seed):
    scene = factory()
    n_pairs = len(scene.pairs)
    xpos, xmat, vel = scene.poses(n_envs, seed)
    eng = make_engine(scene, n_envs)
    eng.step(xpos, xmat, vel)
    res, wrench = eng.pair_results().copy(), eng.geom_wrenches().copy()
    assert int(res["n_polygons"].sum()) > 10 * n_envs

    # 1. the same inputs again: bit-identical (fixed-order sums; where a candidate lands in the pool is decided by
    #    atomics, what is added to what is not)
    eng.step(xpos, xmat, vel)
    assert eng.pair_results().tobytes() == res.tobytes() and eng.geom_wrenches().tobytes() == wrench.tobytes()

    # 2. environments exchange nothing: permuting the batch permutes the results, bit for bit
    perm = np.random.default_rng(seed).permutation(n_envs)
    eng.step(xpos[perm], xmat[perm], vel[perm])
    assert eng.pair_results().tobytes() == res[perm].tobytes()
    assert eng.geom_wrenches().tobytes() == wrench[perm].tobytes()

    # 3. action = -reaction: every pair adds (F, tau) to geom M and subtracts the same numbers from geom N
    #    (plugin.cpp:477-482 applies +f and -f at the same point); the per-geom wrenches of an env sum to zero exactly
    #    when each geom is in one pair, and always to rounding
    total = wrench.sum(axis=1)
    scale = np.abs(wrench).max(axis=(1, 2)) + 1e-300
    assert np.all(np.abs(total).max(axis=1) <= 1e-12 * scale)
    acc = np.zeros_like(wrench)
    for p in range(n_pairs):
        for e_idx, sign in (("gM", 1.0), ("gN", -1.0)):
            g = res[e_idx][:, p].astype(np.int64)
            np.add.at(acc, (np.arange(n_envs), g), sign * np.concatenate([res["F"][:, p], res["tau"][:, p]], axis=1))
    assert np.all(np.abs(acc - wrench) <= 1e-12 * scale[:, None, None])

    # 4. EVERY environment of the full batch against the oracle (same bars as the small-batch tests; round 1 sampled 24)
    eng.step(xpos, xmat, vel)  # the emitted lists on the device are those of the last step: back to the original order
    orc = make_oracle(scene)
    worst = 0.0
    for e in range(n_envs):
        ref_pairs, _ = oracle_env(orc, scene, xpos[e], xmat[e], vel[e], sensors=False)
        worst = max(worst, compare_env(res[e], [eng.emitted(e, p) for p in range(n_pairs)], ref_pairs))
        for g in range(scene.n_geoms):
            worst = max(worst, compare_wrench(wrench[e, g], orc.geom_wrench(g)))
    print("full batch %s: %d envs against the oracle, worst relative error %.2e" % (scene.name, n_envs, worst))
    eng.close()


def test_full_size_rigid_motion_invariance(hcs_lib):
    """Moving the whole world by one rigid motion rotates every force and leaves areas and counts alone (the engine
    works in the soft geom's frame, so this exercises the pose algebra rather than the clipper). Half-space normals
    and friction directions rotate with the world; torques are about the world origin: tau' = Q tau + t x (Q F)."""
    scene = scenes.sphere_on_box()
    n_envs = 4096
    xpos, xmat, vel = scene.poses(n_envs, 99)
    Q, t = _random_rigid_motion(np.random.default_rng(11))
    R = xmat.reshape(n_envs, -1, 3, 3)
    xpos2 = xpos.reshape(n_envs, -1, 3) @ Q.T + t
    xmat2 = (Q @ R).reshape(xmat.shape)
    v6 = vel.reshape(n_envs, -1, 2, 3) @ Q.T  # [omega, v] of the geom origin, world aligned (plugin.cpp:107-114)
    eng = make_engine(scene, n_envs)
    eng.step(xpos, xmat, vel)
    a = eng.pair_results().copy()
    eng.step(xpos2.reshape(xpos.shape), xmat2, v6.reshape(vel.shape))
    b = eng.pair_results().copy()
    eng.close()
    assert np.array_equal(a["n_polygons"], b["n_polygons"]) and np.array_equal(a["n_faces"], b["n_faces"])
    hit = a["n_polygons"][:, 0] > 0
    assert hit.sum() > n_envs // 2
    F, Fq = a["F"][hit, 0] @ Q.T, b["F"][hit, 0]
    fs = np.linalg.norm(F, axis=1)
    assert np.all(np.linalg.norm(Fq - F, axis=1) <= 1e-8 * fs)
    tau = a["tau"][hit, 0] @ Q.T + np.cross(t, F)
    assert np.all(np.linalg.norm(b["tau"][hit, 0] - tau, axis=1) <= 1e-8 * np.maximum(np.linalg.norm(tau, axis=1), 0.1 * fs))
    assert np.all(np.abs(a["area"][hit, 0] - b["area"][hit, 0]) <= 1e-8 * a["area"][hit, 0])
    cen = a["centroid"][hit, 0] @ Q.T + t
    assert np.all(np.linalg.norm(b["centroid"][hit, 0] - cen, axis=1) <= 1e-8)


@pytest.mark.parametrize("name,triangle", [("sphere_on_box", False), ("sphere_on_box", True), ("soft_soft", False),
                                           ("objects_on_plane", False), ("objects_on_plane", True), ("myrmex", True)])
def test_face_vertices_are_the_faces_the_reference_outlines(hcs_lib, name, triangle):
    """f4 (plugin.cpp:509-516, 525-555): visualizeMeshElement walks the vertices of face pc.face of the surface's
    mesh_W.  The per-face dump with hcs_config.face_vertices returns those vertices: same faces, same vertices, same
    winding as the oracle's surface (kPolygon polygons, kTriangle fan triangles; M/N-swapped pairs reversed the way
    SwapMAndN does), and they are consistent with the PointCollision of the face (centroid, normal)."""
    scene = {"sphere_on_box": scenes.sphere_on_box, "soft_soft": scenes.soft_soft,
             "objects_on_plane": scenes.objects_on_plane, "myrmex": lambda: scenes.myrmex("box", 8)}[name]()
    scene.triangle = triangle
    n_envs = 3
    eng = make_engine(scene, n_envs, max_faces=1 << 16, face_vertices=True)
    orc = make_oracle(scene)
    xpos, xmat, vel = scene.poses(n_envs, 99)
    eng.step(xpos, xmat, vel)
    faces, verts = eng.faces(), eng.face_vertices()
    assert len(faces) == len(verts) > 0
    n_checked = 0
    for e in range(n_envs):
        orc.step(xpos[e], xmat[e], vel[e])
        for p in range(len(scene.pairs)):
            ref_nv, ref_v = orc.pair_face_vertices(p)
            ref_pc = orc.pair_faces(p)
            sel = np.flatnonzero((faces["env"] == e) & (faces["pair"] == p))
            assert len(sel) == len(ref_nv)
            if not len(sel):
                continue
            # match faces by their quadrature point (distinct per face)
            key_g = np.lexsort(np.round(faces["p"][sel], 9).T[::-1])
            key_o = np.lexsort(np.round(ref_pc[:, 0:3], 9).T[::-1])
            for ig, io in zip(sel[key_g], key_o):
                nv = 3 if scene.triangle else int(faces["nverts"][ig])
                assert nv == ref_nv[io]
                assert np.allclose(faces["p"][ig], ref_pc[io, 0:3], rtol=0, atol=1e-12)
                assert np.allclose(verts[ig, :nv], ref_v[io, :nv], rtol=0, atol=1e-12), (e, p, verts[ig, :nv], ref_v[io, :nv])
                assert np.all(verts[ig, nv:] == 0)
                # the outline is wound counter-clockwise about the PointCollision normal
                v = verts[ig, :nv]
                nrm = sum(np.cross(v[i] - v[0], v[i + 1] - v[0]) for i in range(1, nv - 1))
                assert np.dot(nrm, faces["n"][ig]) > 0.999999 * np.linalg.norm(nrm)
                n_checked += 1
    assert n_checked > 10
    eng.close()


def test_face_vertices_need_the_config_flag(hcs_lib):
    from mujoco_contact_surfaces_b200.engine import HcsError
    eng = make_engine(scenes.sphere_on_box(), 1, max_faces=1024)
    eng.step(*scenes.sphere_on_box().poses(1, 1))
    with pytest.raises(HcsError):
        eng.face_vertices()
    eng.close()


def _scaled_sphere_on_box(scale, offset):
    """C1 with every length multiplied by `scale` and both geoms moved `offset` metres away from the world origin."""
    S = scenes
    geoms = [S.Geom("box0", S.GEOM_BOX, [0.1 * scale] * 3, [0, 1.0, 0.1 * scale, 0.3, 0.3]),
             S.Geom("sphere0", S.GEOM_SPHERE, [0.08 * scale], [5e4, 5.0, 0.05 * scale, 0.3, 0.3])]
    sc = S.Scene("scaled_sphere_on_box", geoms, [(1, 0)], triangle=False)

    def pose(rng, env, xpos, xmat, vel):
        d = [1e-6, 0.001, 0.012, 0.03][env % 4] * scale
        dx, dy = rng.uniform(-0.05, 0.05, size=2) * scale
        xpos[0] = [offset, -offset, 0.1 * scale]
        xmat[0] = np.eye(3).reshape(-1)
        xpos[1] = [offset + dx, -offset + dy, (0.2 + 0.08) * scale - d]
        xmat[1] = (np.eye(3) if env % 8 == 0 else S.random_rotation(rng)).reshape(-1)
        vel[1] = S.random_velocity(rng, 0.1 * scale, 1.0)

    sc.pose_fn = pose
    return sc


def _near_equal_soft_spheres(sep):
    """Two soft spheres of the same size and modulus whose centres are `sep` apart: where they overlap deeply the two
    pressure gradients nearly cancel, the regime in which the float filter of the soft-soft leaf test must stand back."""
    S = scenes
    geoms = [S.Geom("a", S.GEOM_SPHERE, [0.05], [5e4, 5.0, 0.025, 0.3, 0.3]),
             S.Geom("b", S.GEOM_SPHERE, [0.05], [5e4, 5.0, 0.025, 0.3, 0.3])]
    sc = S.Scene("near_equal_soft_spheres", geoms, [(0, 1)], triangle=False)

    def pose(rng, env, xpos, xmat, vel):
        d = rng.normal(size=3)
        xpos[0] = [0.3, -0.2, 0.1]
        xpos[1] = xpos[0] + d / np.linalg.norm(d) * sep
        xmat[0] = S.random_rotation(rng).reshape(-1)
        xmat[1] = (xmat[0].reshape(3, 3) @ S.rot_zyx(*(rng.normal(size=3) * (1e-7 if env % 2 else 0.3)))).reshape(-1)
        vel[0], vel[1] = S.random_velocity(rng, 0.1, 1.0), S.random_velocity(rng, 0.1, 1.0)

    sc.pose_fn = pose
    return sc


@pytest.mark.parametrize("scene_fn", [lambda: _scaled_sphere_on_box(1.0, 0.0), lambda: _scaled_sphere_on_box(1e3, 0.0),
                                      lambda: _scaled_sphere_on_box(1e-2, 0.0), lambda: _scaled_sphere_on_box(1.0, 1e3),
                                      lambda: _near_equal_soft_spheres(0.06), lambda: _near_equal_soft_spheres(1e-3),
                                      lambda: _near_equal_soft_spheres(1e-7)],
                         ids=["unit", "x1000", "x0.01", "1km_from_origin", "spheres_6cm", "spheres_1mm", "spheres_100nm"])
def test_float_leaf_filters_do_not_change_results(hcs_lib, scene_fn):
    """The broadphase's float leaf filters (kernels_broadphase.cu, HCS_BP_LEAF32) may only reject pairs the exact
    early-outs reject: emitted sets and wrenches must match the oracle whatever the length scale (their margins are
    relative to the largest coordinate involved), for grazing contacts (1e-6 x size deep), axis-aligned poses, and
    when the two gradients of a soft-soft pair nearly cancel."""
    scene = scene_fn()
    # nearly coincident spheres: almost every tet pair whose boxes overlap survives (the plane is ill-defined), far
    # more than the default candidate pool expects
    kw = dict(max_candidates_per_slice=100000) if scene.name == "near_equal_soft_spheres" else {}
    worst = _run_scene(scene, 8 if kw else 32, seed=5, hcs_lib=hcs_lib, **kw)
    assert worst < 1e-8


def test_flat_sensor_reconfigure(hcs_lib):
    """FlatTactileSensor::dynamicParamCallback (flat_tactile_sensor.cpp:48-125): sampling_resolution, window and sigma of
    an existing sensor change at run time; the next image equals the one of a sensor that was loaded that way."""
    n_envs = 3
    scene = scenes.myrmex("box", sampling_resolution=4)
    eng = make_engine(scene, n_envs)
    xpos, xmat, vel = scene.poses(n_envs, seed=7)
    eng.step(xpos, xmat, vel, with_sensors=True)
    before = eng.sensor_image(0).copy()
    eng.update_flat_sensor(0, 8, 1, 0.1)  # 8 x 8 rays per taxel, gauss window
    eng.step(xpos, xmat, vel, with_sensors=True)
    after = eng.sensor_image(0)
    target = scenes.myrmex("box", sampling_resolution=8, window=1, sigma=0.1)
    orc = make_oracle(target)
    for e in range(n_envs):
        _, imgs = oracle_env(orc, target, xpos[e], xmat[e], vel[e], sensors=True)
        err, nbad = compare_images(after[e], imgs[0])
        assert nbad == 0, "reconfigured image: %d taxels beyond %.0e (max rel err %.3e)" % (nbad, TAXEL_RTOL, err)
    assert not np.array_equal(before, after)
    eng.close()


# ---- reference-pinned: the CUDA raster against the REFERENCE's compiled ray caster --------------------------------------
def _ref_cases():
    import os
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_bvh_vectors.npz")) as z:
        return [c.split(":") for c in z["cases"]], {k: z[k] for k in z.files if k.endswith(("_image", "_meta"))}


def _ref_scene(presser, resolution, S):
    if presser == "multi":
        return scenes.myrmex_multi(sampling_resolution=S, resolution=resolution)
    return scenes.myrmex(presser, S, resolution=resolution)


@pytest.mark.parametrize("name,presser", _ref_cases()[0])
def test_cuda_raster_matches_the_reference_ray_caster_vectors(hcs_lib, name, presser):
    """tests/golden/ref_bvh_vectors.npz holds taxel images made with the reference's own compiled BVH/TLAS
    (oracle/_ref, scripts/make_ref_golden.py) inside the flat-sensor loop, at the corners of the benchmark grid of
    benchmark_flat.cpp:282-373 plus a soft presser and a three-surface TLAS case.  1e-6 per taxel, no floor."""
    gold = _ref_cases()[1]
    res, S, seed, env = gold[name + "_meta"]
    S, seed, env = int(S), int(seed), int(env)
    scene = _ref_scene(presser, float(res), S)
    xpos, xmat, vel = scene.poses(env + 1, seed=seed)
    eng = make_engine(scene, env + 1)
    eng.step(xpos, xmat, vel, with_sensors=True)
    img = eng.sensor_image(0)[env]
    eng.close()
    ref = gold[name + "_image"]
    assert (ref > 0).sum() >= 3
    err, nbad = compare_images(img, ref)
    assert nbad == 0, "taxel image vs reference ray caster: %d taxels beyond %.0e (max rel err %.3e)" % (nbad, TAXEL_RTOL, err)


@pytest.mark.parametrize("sse", [False, True])
@pytest.mark.parametrize("presser,resolution,S", [("box", 0.025, 20), ("plate", 0.025, 16), ("spot", 0.025, 8), ("box", 0.0025, 4),
                                                   ("multi", 0.025, 8)])
def test_cuda_raster_matches_the_live_reference_ray_caster(hcs_lib, presser, resolution, S, sse):
    """The same against oracle/_ref live (the prebuilt libraries travel to the GPU box), fresh seeds, both builds."""
    from oracle import oracle as O
    if not O.ref_available(sse):
        pytest.skip("oracle/_ref not present")
    O.use_reference_caster(sse)
    scene = _ref_scene(presser, resolution, S)
    n_envs = 3
    xpos, xmat, vel = scene.poses(n_envs, seed=515)
    eng, orc = make_engine(scene, n_envs), make_oracle(scene)
    eng.step(xpos, xmat, vel, with_sensors=True)
    imgs = eng.sensor_image(0)
    lit = 0
    for e in range(n_envs):
        orc.step(xpos[e], xmat[e], vel[e])
        ref = orc.sensor_image(0, use_bvh=2)
        lit += int((ref > 0).sum())
        err, nbad = compare_images(imgs[e], ref)
        assert nbad == 0, "env %d: %d taxels beyond %.0e (max rel err %.3e)" % (e, nbad, TAXEL_RTOL, err)
    assert lit > 0
    eng.close()


@pytest.mark.parametrize("name,n_envs", [("sphere_on_box", 300), ("objects_on_plane", 64), ("myrmex_box", 6)])
def test_pipelined_steps_equal_synchronous_steps(hcs_lib, name, n_envs):
    """hcs_step_async / hcs_wait: a run of pipelined steps on different pose sets (two in flight, results in caller-owned
    buffers, inputs reused as the contract allows) gives bit for bit what hcs_step gives for each set."""
    import ctypes as C
    factory = {"sphere_on_box": scenes.sphere_on_box, "objects_on_plane": scenes.objects_on_plane,
               "myrmex_box": lambda: scenes.myrmex("box", 4)}[name]
    scene = factory()
    with_sensors = bool(scene.sensors)
    eng = make_engine(scene, n_envs)
    sets = [[np.ascontiguousarray(a) for a in scene.poses(n_envs, seed=70 + i)] for i in range(5)]
    ref_w, ref_img = [], []
    for xp, xm, ve in sets:
        eng.step(xp, xm, ve, with_sensors=with_sensors)
        ref_w.append(eng.geom_wrenches().copy())
        ref_img.append(eng.sensor_image(0).copy() if with_sensors else None)
    outs = [np.full((n_envs, scene.n_geoms, 6), np.nan) for _ in sets]
    imgs = [np.full_like(ref_img[0], np.nan) if with_sensors else None for _ in sets]
    tickets = []
    for i, (xp, xm, ve) in enumerate(sets):
        t = eng.step_async(xp.ctypes.data, xm.ctypes.data, ve.ctypes.data, with_sensors, outs[i].ctypes.data,
                           [imgs[i].ctypes.data] if with_sensors else None)
        tickets.append(t)
        if i >= 1:
            eng.wait(tickets[i - 1])
            assert outs[i - 1].tobytes() == ref_w[i - 1].tobytes()
    eng.wait(tickets[-1])
    for i in range(len(sets)):
        assert outs[i].tobytes() == ref_w[i].tobytes()
        if with_sensors:
            assert imgs[i].tobytes() == ref_img[i].tobytes()
    with pytest.raises(Exception):
        eng.wait(tickets[0])  # its slot has been reused
    eng.close()


def _cuda_device_count():
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0:
            return n.value
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_envs,n_blocks", [("sphere_on_box", 37, 2), ("objects_on_plane", 16, 3), ("myrmex_box", 5, 2)])
def test_multi_device_context_equals_one_context(hcs_lib, name, n_envs, n_blocks):
    """hcs_multi (one context spanning GPUs, env blocks by index, one C++ host thread per block): per-geom wrenches,
    per-pair results and taxel images are bit for bit those of a single context with the same environments, through the
    synchronous and the pipelined step.  On a box with several GPUs the blocks sit on different devices (SURVEY.md section
    4 item 5: multi-GPU bit-identity on hardware), with one GPU they share it."""
    from mujoco_contact_surfaces_b200 import MultiDeviceEngine, REP_POLYGON, REP_TRIANGLE
    factory = {"sphere_on_box": scenes.sphere_on_box, "objects_on_plane": scenes.objects_on_plane,
               "myrmex_box": lambda: scenes.myrmex("box", 4)}[name]
    scene = factory()
    with_sensors = bool(scene.sensors)
    n_dev = _cuda_device_count()
    devices = [k % max(1, n_dev) for k in range(n_blocks)]
    xp, xm, ve = [np.ascontiguousarray(a) for a in scene.poses(n_envs, seed=91)]
    one = make_engine(scene, n_envs)
    one.step(xp, xm, ve, with_sensors=with_sensors)
    ref_w, ref_p = one.geom_wrenches().copy(), one.pair_results().copy()
    ref_img = one.sensor_image(0).copy() if with_sensors else None
    one.close()
    multi = MultiDeviceEngine(n_envs, devices, representation=REP_TRIANGLE if scene.triangle else REP_POLYGON,
                              apply_contact_forces=scene.apply_forces, **scene.engine_kwargs(n_envs))
    scenes.configure(multi, scene)
    multi.finalize()
    blocks = multi.blocks()
    assert len(blocks) == n_blocks and sum(c for _, c in blocks) == n_envs and blocks[0][0] == 0
    for _ in range(2):
        multi.step(xp, xm, ve, with_sensors=with_sensors)
        assert multi.geom_wrenches().tobytes() == ref_w.tobytes()
        got = multi.pair_results()
        for f in ("F", "tau", "centroid", "area", "n_polygons", "n_faces", "n_points", "n_clipped"):
            assert got[f].tobytes() == ref_p[f].tobytes(), f
        if with_sensors:
            assert multi.sensor_image(0).tobytes() == ref_img.tobytes()
    out = np.full((n_envs, scene.n_geoms, 6), np.nan)
    img = np.full_like(ref_img, np.nan) if with_sensors else None
    t = multi.step_async(xp.ctypes.data, xm.ctypes.data, ve.ctypes.data, with_sensors, out.ctypes.data,
                         [img.ctypes.data] if with_sensors else None)
    multi.wait(t)
    assert out.tobytes() == ref_w.tobytes()
    if with_sensors:
        assert img.tobytes() == ref_img.tobytes()
    multi.close()


def _resized(scene_fn, sizes_by_geom):
    sc = scene_fn()
    for g, s in sizes_by_geom.items():
        sc.geoms[g].size = np.resize(np.asarray(s, dtype=np.float64), 3)
    return sc


@pytest.mark.gpu
def test_per_env_sizes_equal_one_engine_per_size(hcs_lib):
    """hcs_set_env_sizes (domain-randomised sizes per environment; the reference rebuilds one geom of
    its one mjData in onGeomChanged, plugin.cpp:828-975).  Every environment of a batch has its own sphere radius (vertices
    and pressures generated on the GPU from the unit mesh) and its own rigid box (host generator per environment); fields,
    element records and LBVHs are built on the GPU per environment.  Environment e gives, bit for bit, what a one-environment
    engine configured with e's sizes gives on e's poses, and matches the oracle configured with e's sizes."""
    n_envs = 10
    rng = np.random.default_rng(5)
    radii = rng.uniform(0.07, 0.12, n_envs)     # refinement level 2 for hint 0.05 needs 0.0653 < r <= 0.128
    halves = rng.uniform(0.08, 0.1, (n_envs, 3))  # 2 x 2 x 2 cells for hint 0.1 needs 0.05 < half size <= 0.1
    scene = scenes.sphere_on_box()
    eng = make_engine(scene, n_envs)
    s_sph = np.zeros((n_envs, 3))
    s_sph[:, 0] = radii
    eng.set_env_sizes(1, s_sph)   # after hcs_finalize: the context is rebuilt
    eng.set_env_sizes(0, halves)
    xp, xm, ve = scene.poses(n_envs, seed=17)
    for e in range(n_envs):       # box top at 0.1 + half height, sphere pressed 6 .. 15 mm into it (the level-2 sphere is faceted)
        xp[e, 1, 2] = xp[e, 0, 2] + halves[e, 2] + radii[e] - 0.001 * (6 + e)
    eng.step(xp, xm, ve)
    W, P = eng.geom_wrenches().copy(), eng.pair_results().copy()
    assert (P["n_polygons"][:, 0] > 0).all()
    worst = 0.0
    for e in range(n_envs):
        one_scene = _resized(scenes.sphere_on_box, {0: halves[e], 1: [radii[e]] * 3})
        one = make_engine(one_scene, 1)
        one.step(xp[e:e + 1], xm[e:e + 1], ve[e:e + 1])
        assert one.geom_wrenches().tobytes() == W[e:e + 1].tobytes(), "env %d" % e
        got = one.pair_results()
        for f in ("n_polygons", "n_faces", "n_points", "n_clipped"):
            assert got[f][0, 0] == P[f][e, 0]
        one.close()
        if e < 4:
            orc = make_oracle(one_scene)
            ref_pairs, _ = oracle_env(orc, one_scene, xp[e], xm[e], ve[e], sensors=False)
            worst = max(worst, compare_env(P[e], [eng.emitted(e, 0)], ref_pairs))
    assert worst < 1e-8
    # back to one size: the plain configuration again
    eng.set_env_sizes(1, None)
    eng.set_env_sizes(0, None)
    eng.step(*scene.poses(n_envs, seed=17))
    ref = make_engine(scenes.sphere_on_box(), n_envs)
    ref.step(*scene.poses(n_envs, seed=17))
    assert eng.geom_wrenches().tobytes() == ref.geom_wrenches().tobytes()
    ref.close()
    # a radius that needs another refinement level is refused, with the environment named
    s_bad = s_sph.copy()
    s_bad[3, 0] = 0.03
    with pytest.raises(Exception, match="environment 3"):
        eng.set_env_sizes(1, s_bad)
    eng.close()


@pytest.mark.gpu
def test_per_env_sizes_soft_soft_and_half_space(hcs_lib):
    """Per-environment ellipsoid semi-axes on both sides of a soft-soft pair (equal-pressure plane, per-environment LBVH on the
    tree side and per-environment query tets) and per-environment soft geoms on a rigid half space."""
    n_envs = 6
    rng = np.random.default_rng(9)
    f = rng.uniform(0.9, 1.0, (n_envs, 1))  # (ell0 changes its refinement level at 1.02 x, ell1 at 0.85 x)
    a, b = np.array([0.05, 0.04, 0.03]) * f, np.array([0.04, 0.04, 0.06]) * f[::-1]
    scene = scenes.soft_soft()
    eng = make_engine(scene, n_envs)
    eng.set_env_sizes(0, a)
    eng.set_env_sizes(1, b)
    xp, xm, ve = scene.poses(n_envs, seed=23)
    eng.step(xp, xm, ve)
    W = eng.geom_wrenches().copy()
    hit = 0
    for e in range(n_envs):
        one = make_engine(_resized(scenes.soft_soft, {0: a[e], 1: b[e]}), 1)
        one.step(xp[e:e + 1], xm[e:e + 1], ve[e:e + 1])
        assert one.geom_wrenches().tobytes() == W[e:e + 1].tobytes(), "env %d" % e
        hit += int(one.pair_results()["n_polygons"][0, 0] > 0)
        one.close()
    assert hit > 0
    eng.close()
    # half space: spheres of per-environment radius on the plane
    scene = scenes.objects_on_plane()
    eng = make_engine(scene, n_envs)
    radii = rng.uniform(0.045, 0.055, n_envs)
    s = np.zeros((n_envs, 3))
    s[:, 0] = radii
    eng.set_env_sizes(1, s)
    xp, xm, ve = scene.poses(n_envs, seed=29)
    eng.step(xp, xm, ve)
    W = eng.geom_wrenches().copy()
    for e in range(n_envs):
        one = make_engine(_resized(scenes.objects_on_plane, {1: [radii[e]] * 3}), 1)
        one.step(xp[e:e + 1], xm[e:e + 1], ve[e:e + 1])
        assert one.geom_wrenches().tobytes() == W[e:e + 1].tobytes(), "env %d" % e
        one.close()
    eng.close()


@pytest.mark.gpu
def test_gpu_sphere_generation_equals_host_generator(hcs_lib, monkeypatch):
    """K0 (kernels_meshgen.cu): sphere and ellipsoid meshes are generated on the GPU - topology by sorting the edge slots of
    every refinement level into first-use order, vertices, pressures - and must equal the host generator's mesh bit for
    bit (HCS_MESHGEN_CHECK makes hcs_finalize compare them).  Levels 0 .. 7 (131 072 tets, the pads of config 5), soft and
    rigid, sphere and ellipsoid; the device mesh is also the one the oracle builds (vertex order included)."""
    from mujoco_contact_surfaces_b200 import HydroelasticEngine, GEOM_ELLIPSOID, GEOM_SPHERE
    from oracle.oracle import OracleScene
    monkeypatch.setenv("HCS_MESHGEN_CHECK", "1")
    eng = HydroelasticEngine(1)
    orc = OracleScene(False, True)
    cases = []
    for level, hint in ((0, 0.2), (1, 0.08), (2, 0.05), (3, 0.02), (5, 0.005)):
        cases.append((GEOM_SPHERE, [0.08, 0, 0], [5e4, 5.0, hint, 0.3, 0.3]))   # soft spheres
    cases.append((GEOM_SPHERE, [0.05, 0, 0], [0, 1.0, 0.02, 0.5, 0.5]))         # rigid sphere
    cases.append((GEOM_ELLIPSOID, [0.05, 0.04, 0.03], [5e4, 5.0, 0.01, 0.3, 0.3]))
    cases.append((GEOM_ELLIPSOID, [0.05, 0.03, 0.04], [0, 1.0, 0.02, 0.5, 0.5]))  # rigid ellipsoid
    cases.append((GEOM_ELLIPSOID, [0.010, 0.010, 0.012], [1e5, 2.0, 0.00025, 0.6, 0.6]))  # a fingertip pad of config 5: level 7
    for t, size, props in cases:
        eng.add_geom(t, size, props)
        orc.add_geom(t, size, props)
    eng.set_pairs([(0, 5)])
    eng.finalize()  # raises if any GPU-generated mesh differs from the host generator's
    counts = []
    for g in range(len(cases)):
        m, o = eng.geom_mesh(g), orc.geom_mesh(g)
        for key in o:
            if key != "kind":
                assert np.array_equal(m[key], o[key]), "geom %d: %s differs" % (g, key)
        counts.append(len(m["elems"]))
    assert counts[0] == 8 and counts[2] == 128 and counts[-1] == 131072
    eng.close()


@pytest.mark.gpu
def test_refinalize_with_more_pairs_and_empty_configurations(hcs_lib):
    """hcs_set_pairs + hcs_finalize on a running context (what the adapter does when MuJoCo reports a geom pair for the
    first time, plugin.cpp:255-318) keeps the geoms' device records and gives the old pairs the same results; a context
    without pairs, and a multi-device context with more blocks than environments, step to exact zeros / the same bytes."""
    from mujoco_contact_surfaces_b200 import MultiDeviceEngine, REP_POLYGON
    scene = scenes.objects_on_plane()
    n_envs = 5
    xp, xm, ve = scene.poses(n_envs, seed=41)
    full = make_engine(scene, n_envs)
    full.step(xp, xm, ve)
    ref = full.pair_results().copy()
    eng = make_engine(scenes.objects_on_plane(), n_envs)
    eng.set_pairs([])                      # no pairs at all: a legal configuration
    eng.finalize()
    eng.step(xp, xm, ve)
    assert not eng.geom_wrenches().any() and eng.pair_results().shape == (n_envs, 0)
    eng.set_pairs(scene.pairs[:2])         # pairs appear one after the other
    eng.finalize()
    eng.step(xp, xm, ve)
    first = eng.pair_results().copy()
    eng.set_pairs(scene.pairs)
    eng.finalize()
    eng.step(xp, xm, ve)
    got = eng.pair_results()
    for f in ("F", "tau", "centroid", "area", "n_polygons", "n_faces", "n_points"):
        assert got[f].tobytes() == ref[f].tobytes(), f
        assert first[f].tobytes() == ref[f][:, :2].tobytes(), f
    assert eng.geom_wrenches().tobytes() == full.geom_wrenches().tobytes()
    eng.close()
    # three blocks for two environments: one block stays empty
    multi = MultiDeviceEngine(2, [0, 0, 0], representation=REP_POLYGON, apply_contact_forces=scene.apply_forces)
    scenes.configure(multi, scene)
    multi.finalize()
    assert [c for _, c in multi.blocks()] == [1, 1, 0]
    multi.step(xp[:2], xm[:2], ve[:2])
    assert multi.geom_wrenches().tobytes() == full.geom_wrenches()[:2].tobytes()
    multi.close()
    full.close()


@pytest.mark.gpu
def test_kernel_shared_memory_opt_in_survives_a_change_of_scene(hcs_lib):
    """Regression (found by the multi-device test on a 2-GPU box): the dynamic shared memory attribute of a kernel is a LIMIT.
    The launcher used to set it to every launch's size, with one bookkeeping array per function-pointer type, which the two
    finalize kernels share: a scene with three pairs (limit of finalize_kernel := 3024 B), one with two geoms (the shared
    entry := 3072 B, set for finalize_env_kernel) and then one with four pairs (3072 B: "already set") made that last launch
    fail with "invalid argument".  Runs in a fresh interpreter, because the order of the first launches is the point."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, "tests")
import os

import numpy as np
from mujoco_contact_surfaces_b200 import scenes
from parity_utils import make_engine
for factory, n in ((scenes.myrmex_multi, 3), (scenes.sphere_on_box, 5), (scenes.objects_on_plane, 16)):
    sc = factory()
    xp, xm, ve = [np.ascontiguousarray(a) for a in sc.poses(n, seed=3)]
    e = make_engine(sc, n)
    for _ in range(3):
        e.step(xp, xm, ve, with_sensors=bool(sc.sensors))
    assert np.isfinite(e.geom_wrenches()).all()
    e.close()
print("ok")
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_soft_box_stack_meets_the_closed_form_on_the_gpu(hcs_lib):
    """The soft-soft path of the CUDA engine against an ANALYTIC answer (not only against the oracle): two medial-axis boxes
    stacked with overlap d, the upper one wider; F_z = E/c (4ab t - 2(a+b) t^2 + 4/3 t^3), t = d/2, and the surface area
    (tests/test_oracle_kat.py::soft_box_stack_closed_form).  A batch of rigid placements and overlaps, one per environment."""
    from mujoco_contact_surfaces_b200 import HydroelasticEngine, GEOM_BOX, REP_POLYGON, REP_TRIANGLE
    from test_oracle_kat import rot, soft_box_stack_closed_form
    E, c = 5e4, 0.02
    overlaps = [0.004, 0.01, 0.016, 0.007, 0.012, 0.002]
    n = len(overlaps)
    rng = np.random.default_rng(11)
    # (lower box, upper box, footprint a x b, area known): upper wider both ways; upper narrower in x and wider in y (crossed:
    # roof pieces of both fields and vertical pieces in the corners, test_crossed_soft_boxes_force_closed_form)
    for lo, up, fa, fb, with_area, rep in (((0.06, 0.04, c), (0.12, 0.10, c), 0.06, 0.04, True, REP_POLYGON),
                                           ((0.08, 0.04, c), (0.03, 0.10, c), 0.03, 0.04, False, REP_POLYGON),
                                           ((0.06, 0.04, c), (0.12, 0.10, c), 0.06, 0.04, True, REP_TRIANGLE)):
        eng = HydroelasticEngine(n, representation=rep)
        eng.add_geom(GEOM_BOX, list(lo), [E, 0, 0, 0.3, 0.3])
        eng.add_geom(GEOM_BOX, list(up), [E, 0, 0, 0.3, 0.3])
        eng.set_pairs([(0, 1)])
        eng.finalize()
        xp, xm, ve = np.zeros((n, 2, 3)), np.zeros((n, 2, 9)), np.zeros((n, 2, 6))
        Rs = []
        for e, d in enumerate(overlaps):
            R = np.eye(3) if e == 0 else rot(rng.normal(size=3), rng.uniform(0, np.pi))
            p = rng.uniform(-0.5, 0.5, size=3)
            xp[e, 0], xp[e, 1] = p, p + R @ np.array([0.003, -0.002, 2 * c - d])
            xm[e, 0] = xm[e, 1] = R.reshape(-1)
            Rs.append(R)
        eng.step(xp, xm, ve)
        res = eng.pair_results()
        for e, d in enumerate(overlaps):
            force, area = soft_box_stack_closed_form(E, fa, fb, c, d)
            F_local = Rs[e].T @ res["F"][e, 0]
            assert np.isclose(abs(F_local[2]), force, rtol=1e-11), (e, F_local, force)
            assert abs(F_local[0]) < 1e-11 * force and abs(F_local[1]) < 1e-11 * force
            if with_area:
                assert np.isclose(res["area"][e, 0], area, rtol=1e-11)
        eng.close()


@pytest.mark.gpu
def test_faces_without_pressure_gradient_along_their_normal_carry_no_force_on_the_gpu(hcs_lib):
    """plugin.cpp:366-373 on the half-space path of the CUDA engine, against the analytic answer
    (tests/test_oracle_kat.py::test_faces_without_pressure_gradient_along_their_normal_carry_no_force): axis-aligned soft box
    pressed into a rigid plane, F = E d / h (2a - 2d)(2b - 2d), surface area 4ab."""
    from mujoco_contact_surfaces_b200 import HydroelasticEngine, GEOM_BOX, GEOM_PLANE
    E, a, b, c = 5e4, 0.06, 0.04, 0.03
    depths = [0.003, 0.008, 0.015]
    n = len(depths)
    eng = HydroelasticEngine(n)
    eng.add_geom(GEOM_PLANE, [0, 0, 1], [0, 1, 0, 0.3, 0.3])
    eng.add_geom(GEOM_BOX, [a, b, c], [E, 0, 0, 0.3, 0.3])
    eng.set_pairs([(0, 1)])
    eng.finalize()
    xp, xm, ve = np.zeros((n, 2, 3)), np.zeros((n, 2, 9)), np.zeros((n, 2, 6))
    xm[:, :] = np.eye(3).reshape(-1)
    for e, d in enumerate(depths):
        xp[e, 1] = [0.01, 0.02, c - d]
    eng.step(xp, xm, ve)
    res = eng.pair_results()
    for e, d in enumerate(depths):
        assert np.isclose(res["area"][e, 0], 4 * a * b, rtol=1e-12)
        assert np.isclose(abs(res["F"][e, 0][2]), E * d / min(a, b, c) * (2 * a - 2 * d) * (2 * b - 2 * d), rtol=1e-11)
        assert 0 < res["n_points"][e, 0] < res["n_faces"][e, 0]
    eng.close()
