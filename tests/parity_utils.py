"""Shared helpers of the GPU-vs-oracle parity tests (test infrastructure)."""
import numpy as np

from mujoco_contact_surfaces_b200 import scenes
from oracle.oracle import OracleScene

# north-star tolerances (BASELINE.json): candidate set + polygon vertex counts bit-exact,
# per-pair force/torque 1e-8 relative (fp64 geometry mode), taxel images 1e-6 relative.
FORCE_RTOL = 1e-8
TAXEL_RTOL = 1e-6


def make_oracle(scene):
    o = OracleScene(scene.triangle, scene.apply_forces)
    scenes.configure(o, scene)
    return o


def make_engine(scene, n_envs, **kw):
    from mujoco_contact_surfaces_b200 import HydroelasticEngine, REP_POLYGON, REP_TRIANGLE
    kw = dict(scene.engine_kwargs(n_envs), **kw)
    e = HydroelasticEngine(n_envs, representation=REP_TRIANGLE if scene.triangle else REP_POLYGON,
                           apply_contact_forces=scene.apply_forces, **kw)
    scenes.configure(e, scene)
    e.finalize()
    return e


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.linalg.norm(b), floor)
    return np.linalg.norm(a - b) / scale if scale > 0 else np.linalg.norm(a - b)


def oracle_env(o, scene, xpos, xmat, vel, use_bvh=True, sensors=True):
    """Run one env through the oracle; returns per-pair dicts (+ emitted sets) and sensor images."""
    o.step(xpos, xmat, vel, use_bvh=use_bvh)
    pairs = []
    for p in range(len(scene.pairs)):
        r = o.pair_result(p)
        r["emitted"] = set(map(tuple, o.pair_emitted(p).tolist()))
        pairs.append(r)
    images = [o.sensor_image(s) for s in range(len(scene.sensors))] if sensors else []
    return pairs, images


def compare_env(gpu_pairs_row, gpu_emitted, oracle_pairs):
    """Assert the three north-star parity bars for one env. gpu_pairs_row: structured array [n_pairs]."""
    worst = 0.0
    for p, ref in enumerate(oracle_pairs):
        g = gpu_pairs_row[p]
        em = set(map(tuple, gpu_emitted[p].tolist()))
        assert em == ref["emitted"], "emitted (elemM, elemN, nverts) set differs for pair %d: gpu-only %s, oracle-only %s" % (
            p, sorted(em - ref["emitted"])[:5], sorted(ref["emitted"] - em)[:5])
        assert int(g["n_polygons"]) == ref["n_polygons"]
        if not ref["has_surface"]:
            assert np.all(g["F"] == 0) and np.all(g["tau"] == 0)
            continue
        assert (int(g["gM"]), int(g["gN"])) == (ref["gM"], ref["gN"])
        assert int(g["n_faces"]) == ref["n_faces"]
        assert int(g["n_points"]) == ref["n_points"], (int(g["n_points"]), ref["n_points"])
        fscale = np.linalg.norm(ref["F"])
        ef = rel_err(g["F"], ref["F"])
        # torque about the world origin, relative to the torque itself (round 1 also accepted 0.1 |F| as the scale).
        # A torque that cancels to below 1e-6 of |F| x |centroid| carries no more than ~1e-10 of relative information
        # in fp64, so that (never observed on the test scenes) is the only floor left.
        tscale = max(np.linalg.norm(ref["tau"]), 1e-6 * fscale * np.linalg.norm(ref["centroid"]), 1e-300)
        et = np.linalg.norm(g["tau"] - ref["tau"]) / tscale
        ea = abs(g["area"] - ref["area"]) / max(ref["area"], 1e-300)
        ec = np.linalg.norm(g["centroid"] - ref["centroid"]) / max(np.linalg.norm(ref["centroid"]), 1e-3)
        if fscale > 0:
            assert ef < FORCE_RTOL, "force rel err %.3e (pair %d)" % (ef, p)
            assert et < FORCE_RTOL, "torque rel err %.3e (pair %d)" % (et, p)
        assert ea < FORCE_RTOL and ec < FORCE_RTOL, (ea, ec)
        worst = max(worst, ef, et, ea, ec)
    return worst


def compare_images(gpu_img, ref_img, rtol=TAXEL_RTOL):
    """Taxel image parity, element-wise relative to each taxel's own value (no floor: round 1 measured small taxels
    against 1e-3 of the image peak).  Taxels the reference leaves at zero must be exactly zero."""
    gpu_img, ref_img = np.asarray(gpu_img, dtype=np.float64), np.asarray(ref_img, dtype=np.float64)
    zero = ref_img == 0
    if np.any(gpu_img[zero] != 0):
        return float("inf"), int((gpu_img[zero] != 0).sum())
    if zero.all():
        return 0.0, 0
    err = np.zeros_like(ref_img)
    err[~zero] = np.abs(gpu_img[~zero] - ref_img[~zero]) / np.abs(ref_img[~zero])
    return float(err.max()), int((err > rtol).sum())


def compare_wrench(gpu_w, ref_w):
    """Per-geom wrench (F, tau about the world origin): 1e-8 relative to the wrench itself; a geom nothing touched must
    read exactly zero (round 1 let anything below 1e-9 pass)."""
    ref_w = np.asarray(ref_w, dtype=np.float64)
    n = np.linalg.norm(ref_w)
    if n == 0:
        assert np.all(np.asarray(gpu_w) == 0)
        return 0.0
    err = np.linalg.norm(gpu_w - ref_w) / n
    assert err < FORCE_RTOL, "per-geom wrench rel err %.3e" % err
    return err
