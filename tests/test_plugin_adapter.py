"""The C++ host adapter (mujoco_contact_surfaces_b200/plugin) driven like mujoco_ros drives the reference
plugin: load -> collision pass -> passiveCallback, on shim re-creations of the reference's example worlds.
Results are checked against the CPU oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import GEOM_BOX, GEOM_SPHERE, OracleScene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN_DIR = os.path.join(ROOT, "mujoco_contact_surfaces_b200", "plugin")


@pytest.fixture(scope="module")
def plugin_built(hcs_lib):
    subprocess.check_call(["make", "-C", PLUGIN_DIR, "-s"])
    return os.path.join(PLUGIN_DIR, "test_plugin")


def test_adapter_builds_and_mirrors_the_reference_classes(plugin_built):
    """CPU-side: the adapter compiles against the MuJoCo shim and declares the reference's virtual surface."""
    hdr = open(os.path.join(PLUGIN_DIR, "contact_surfaces_plugin.h")).read()
    for name in ("class MujocoContactSurfacesPlugin", "bool load(const mjModel *m, mjData *d) override",
                 "void passiveCallback(const mjModel *model, mjData *data) override",
                 "void renderCallback(const mjModel *model, mjData *data, mjvScene *scene) override",
                 "void onGeomChanged(const mjModel *model, mjData *data, const int geom_id) override",
                 "int collision_cb(const mjModel *m, const mjData *d, mjContact *con, int g1, int g2, mjtNum margin)",
                 "class SurfacePlugin", "bool safe_load(", "void safe_reset()", "struct PointCollision",
                 "struct GeomCollision", "class FlatTactileSensor", "class TactileSensorBase"):
        assert name in hdr, name
    assert os.path.exists(plugin_built)


def _run(binary):
    out = subprocess.run([binary], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    return {d["scenario"]: d for d in map(json.loads, out.stdout.strip().splitlines())}


@pytest.mark.gpu
def test_adapter_sphere_on_box_and_myrmex_match_the_oracle(plugin_built):
    res = _run(plugin_built)
    I3 = np.eye(3).reshape(-1)

    r = res["sphere_on_box"]
    o = OracleScene(triangle_representation=False)
    box = o.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1.0, 0.1, 0.3, 0.3])
    sph = o.add_geom(GEOM_SPHERE, [0.08], [5e4, 5.0, 0.05, 0.3, 0.3])
    o.set_pairs([[sph, box]])
    xpos = np.array([r["box_pos"], r["sphere_pos"]])
    xmat = np.stack([I3, np.array(r["sphere_mat"])])
    vel = np.stack([np.zeros(6), np.array(r["sphere_vel6"])])
    o.step(xpos, xmat, vel)
    q = np.array(r["qfrc_passive"])
    assert q.shape == (12,) and r["vgeoms"] > 0
    for g, dofs in ((box, q[0:6]), (sph, q[6:12])):
        w = o.geom_wrench(g)
        tau_com = w[3:] - np.cross(xpos[g], w[:3])  # mj_applyFT: torque about the body's centre of mass
        assert np.allclose(dofs[:3], w[:3], rtol=1e-8, atol=1e-12)
        assert np.allclose(dofs[3:], tau_com, rtol=1e-8, atol=1e-10)
    assert np.linalg.norm(q[:3]) > 1.0  # there is contact
    # cs::VisualizeSurfaces = 1: every face with a PointCollision is outlined, one connector per edge
    # (plugin.cpp:509-516, 525-555)
    nv, fv = o.pair_face_vertices(0)
    perimeter = sum(np.linalg.norm(fv[i, (k + 1) % n] - fv[i, k]) for i, n in enumerate(nv) for k in range(n))
    assert r["vgeoms"] == r["connectors"] == int(nv.sum()) and len(nv) > 10
    assert abs(r["outline_length"] - perimeter) < 1e-5 * perimeter

    r = res["myrmex_box"]
    o = OracleScene(triangle_representation=True)
    b1 = o.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1.0, 0.05, 0.3, 0.3])
    foam = o.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5.0, 0, 0.3, 0.3])
    o.set_pairs([[b1, foam]])
    o.add_flat_sensor(foam, [0.2, 0.2, 0.02], 0.025, 20)
    xpos = np.array([r["box_pos"], [0, 0, 0.033]])
    o.step(xpos, np.stack([np.array(r["box_mat"]), I3]))
    ref = o.sensor_image(0)
    img = np.array(r["image"], dtype=np.float32)
    # 50 Hz over 45 ms of 1 ms steps: three messages, plus the one after the reconfigure request
    assert (r["cx"], r["cy"]) == (16, 16) and r["publishes"] == 4 and r["fetched_sampling"] == 20
    assert ref.max() > 0
    err = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3 * ref.max())
    assert err.max() < 1e-6
    # dynamic_reconfigure request (flat_tactile_sensor.cpp:48-125): 8 x 8 rays per taxel and a gauss window
    o2 = OracleScene(triangle_representation=True)
    c1 = o2.add_geom(GEOM_BOX, [0.1, 0.1, 0.1], [0, 1.0, 0.05, 0.3, 0.3])
    foam2 = o2.add_geom(GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5.0, 0, 0.3, 0.3])
    o2.set_pairs([[c1, foam2]])
    o2.add_flat_sensor(foam2, [0.2, 0.2, 0.02], 0.025, 8, 1, 0.1)
    o2.step(xpos, np.stack([np.array(r["box_mat"]), I3]))
    ref2 = o2.sensor_image(0)
    img2 = np.array(r["image_reconfigured"], dtype=np.float32)
    assert ref2.max() > 0 and not np.array_equal(img2, img)
    assert (np.abs(img2 - ref2) / np.maximum(np.abs(ref2), 1e-3 * ref2.max())).max() < 1e-6
    w = o.geom_wrench(b1)
    q = np.array(r["qfrc_passive"])
    assert np.allclose(q[:3], w[:3], rtol=1e-8)
    assert np.allclose(q[3:6], w[3:] - np.cross(xpos[0], w[:3]), rtol=1e-8, atol=1e-10)


@pytest.mark.gpu
def test_adapter_curved_sensor_matches_the_oracle(plugin_built):
    """CurvedSensor adapter class on a soft convex-mesh tip: the samples it drew go to the oracle unchanged."""
    r = _run(plugin_built)["curved_tip"]
    I3 = np.eye(3).reshape(-1)
    GEOM_MESH = 7
    o = OracleScene(triangle_representation=True)
    box = o.add_geom(GEOM_BOX, [0.025] * 3, [0, 1.0, 0.01, 0.0, 0.0])
    mv = np.array(r["mesh_vert"], dtype=np.float32).reshape(-1, 3)
    mf = np.array(r["mesh_face"], dtype=np.int32).reshape(-1, 3)
    tip = o.add_geom(GEOM_MESH, [0, 0, 0], [5e4, 5.0, 0.0, 0.0, 0.0], mv, mf)
    o.set_pairs([[box, tip]])
    taxels = np.array(r["taxels"]).reshape(-1, 3)
    cs = o.add_curved_sensor(tip, taxels, np.array(r["normals"]).reshape(-1, 3), np.array(r["sample_pos"]).reshape(-1, 3),
                             np.array(r["sample_nrm"]).reshape(-1, 3), 0.006)
    # (int)0.0008 * area == 0 samples requested: vcglib hands back its whole Monte-Carlo pool (curved_sensor.cpp:280-282)
    assert len(taxels) == 5 and len(r["sample_pos"]) // 3 == 10000 and o.curved_info(cs)[0] > 50
    o.step(np.array([[0, 0, 0.025], r["tip_pos"]]), np.stack([I3, np.array(r["tip_mat"])]))
    assert o.pair_result(0)["n_polygons"] > 0
    ref = o.curved_values(cs)
    val = np.array(r["values"], dtype=np.float32)
    assert r["publishes"] == 1 and ref.max() > 0 and ref[4] == 0  # the taxel on the far side feels nothing
    err = np.abs(val - ref) / np.maximum(np.abs(ref), 1e-3 * ref.max())
    assert err.max() < 1e-6


@pytest.mark.gpu
def test_adapter_taxel_sensor_with_the_fingertip_yaml_keys(plugin_built):
    """TaxelSensor adapter class configured like SENS/config/fingertip.yaml (method squared, sample_method
    area_importance): the published values equal the oracle's."""
    r = _run(plugin_built)["taxel_tip"]
    I3 = np.eye(3).reshape(-1)
    o = OracleScene(triangle_representation=True)
    box = o.add_geom(GEOM_BOX, [0.025] * 3, [0, 1.0, 0.01, 0.0, 0.0])
    mv = np.array(r["mesh_vert"], dtype=np.float32).reshape(-1, 3)
    mf = np.array(r["mesh_face"], dtype=np.int32).reshape(-1, 3)
    tip = o.add_geom(7, [0, 0, 0], [5e4, 5.0, 0.0, 0.0, 0.0], mv, mf)
    o.set_pairs([[box, tip]])
    taxels = np.array([[0.002, 0.0015, -0.005], [-0.002, 0.0015, -0.005], [0.002, -0.0015, -0.005],
                       [-0.002, -0.0015, -0.005], [0, 0, 0.010]])
    ts = o.add_taxel_sensor(tip, taxels, 0.006, 0.001, "squared", False, "area_importance")
    o.step(np.array([[0, 0, 0.025], r["tip_pos"]]), np.stack([I3, np.array(r["tip_mat"])]))
    assert o.pair_result(0)["n_polygons"] > 0
    ref = o.taxel_values(ts)
    val = np.array(r["values"], dtype=np.float32)
    assert r["publishes"] == 1 and ref.max() > 0
    err = np.abs(val - ref) / np.maximum(np.abs(ref), 1e-3 * ref.max())
    assert err.max() < 1e-6
    # visualize: one sphere per taxel, 0.5 mm .. 2.5 mm by pressure / visualize_max_pressure (taxel_sensor.cpp:447-455)
    scale = min(max(float(ref.max()), 0.0), 0.04) / 0.04
    assert r["taxel_markers"] == 5 and abs(r["max_marker_size"] - (0.0005 + scale * 0.002)) < 1e-7


@pytest.mark.gpu
def test_adapter_curved_sensor_poisson_disk_samples(plugin_built):
    """An integer-valued sample_resolution asks for (int)sample_resolution * area samples (curved_sensor.cpp:280): the
    adapter's restatement of vcglib's Poisson-disk sampling keeps every pair of samples at least one disk radius apart."""
    r = _run(plugin_built)["curved_poisson"]
    sample_num = int(3000000 * r["area"])
    radius = np.sqrt(r["area"] / (0.7 * np.pi * sample_num))
    assert sample_num > 1000 and r["min_distance"] >= radius
    assert 0.5 * sample_num < r["n_samples"] < 2.0 * sample_num


@pytest.mark.gpu
def test_batched_adapter_equals_one_plugin_per_mjdata(plugin_built):
    """BatchedContactSurfaces (C++, one hcs_multi context over two blocks) applies to six mjData exactly the generalised
    forces six one-mjData plugin instances apply; the timing scenarios report sane per-step latencies."""
    res = _run(plugin_built)
    b = res["batched"]
    assert b["n"] == 6 and b["blocks"] == 2 and b["pairs"] == 1
    assert b["checksum"] > 0 and b["max_abs_diff"] == 0.0
    for name in ("timing_sphere_on_box", "timing_objects_on_plane"):
        t = res[name]
        assert 0 < t["passive_callback_us"] < 5000 and t["qfrc_checksum"] > 0
