"""Generates tests/golden/steps.npz: seeded inputs and the CPU oracle's outputs for one small batch of each
BASELINE.json config (C1 sphere on box, C2 Myrmex image, C3 soft-soft, C4 objects on a plane).

The reference has no golden vectors for this path ("parity unpinned"); these fixtures do not change
that.  What they pin is THIS repo: the scene generators (poses from the seeds), the oracle (a change in its
arithmetic shows up as a diff of a committed file) and the CUDA path (tests/test_gpu_parity.py::test_golden_steps
compares it with the file, not with a freshly built oracle).  Run from the repo root:

    python tests/golden/make_step_fixtures.py [out.npz]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mujoco_contact_surfaces_b200 import scenes  # noqa: E402
from parity_utils import make_oracle, oracle_env  # noqa: E402

# name -> (scene factory, environments, seed, sensors)
CASES = {
    "c1_sphere_on_box": (lambda: scenes.sphere_on_box(), 8, 1234, False),
    "c1_identity": (lambda: scenes.sphere_on_box(identity_orientation=True), 4, 1234, False),
    "c2_myrmex_box_s4": (lambda: scenes.myrmex("box", 4), 2, 7, True),
    "c3_soft_soft": (lambda: scenes.soft_soft(), 2, 3, False),
    "c3_soft_soft_triangle": (lambda: scenes.soft_soft(triangle=True), 2, 3, False),
    "c4_objects_on_plane": (lambda: scenes.objects_on_plane(), 4, 4096, False),
    "c2b_myrmex_soft_tip_s4": (lambda: scenes.myrmex("soft_tip", 4), 2, 7, True),
    # TaxelSensor with the reference's fingertip.yaml settings (sample_method area_importance) on the Myrmex foam
    "taxel_area_importance": (lambda: scenes.myrmex_taxels("box", "squared", False, "area_importance", 0.002), 2, 31, False),
}
PAIR_FIELDS = ("has_surface", "F", "tau", "area", "centroid", "n_polygons", "n_faces", "n_points", "gM", "gN")


def run_case(name):
    factory, n_envs, seed, sensors = CASES[name]
    scene = factory()
    orc = make_oracle(scene)
    xpos, xmat, vel = scene.poses(n_envs, seed)
    out = {"xpos": xpos, "xmat": xmat, "vel": vel}
    n_pairs = len(scene.pairs)
    acc = {f: [] for f in PAIR_FIELDS}
    wrench = np.zeros((n_envs, scene.n_geoms, 6))
    emitted, emitted_off = [], [0]
    images = [[] for _ in scene.sensors] if sensors else []
    for e in range(n_envs):
        pairs, imgs = oracle_env(orc, scene, xpos[e], xmat[e], vel[e], sensors=sensors)
        for f in PAIR_FIELDS:
            acc[f].append([np.asarray(pairs[p][f]) for p in range(n_pairs)])
        for g in range(scene.n_geoms):
            wrench[e, g] = orc.geom_wrench(g)
        for p in range(n_pairs):
            rows = np.asarray(sorted(pairs[p]["emitted"]), dtype=np.int32).reshape(-1, 3)
            emitted.append(rows)
            emitted_off.append(emitted_off[-1] + len(rows))
        for s, img in enumerate(imgs):
            images[s].append(np.asarray(img, dtype=np.float32))
        for s in range(len(getattr(scene, "taxel_sensors", []))):  # first update: the previous message is all zeros
            out.setdefault("taxel%d" % s, []).append(orc.taxel_values(s))
    for s in range(len(getattr(scene, "taxel_sensors", []))):
        out["taxel%d" % s] = np.asarray(out["taxel%d" % s], dtype=np.float32)
    for f in PAIR_FIELDS:
        out["pair_" + f] = np.asarray(acc[f])
    out["geom_wrench"] = wrench
    out["emitted"] = np.concatenate(emitted) if emitted else np.zeros((0, 3), np.int32)
    out["emitted_off"] = np.asarray(emitted_off, dtype=np.int64)  # [(env, pair)] row ranges, env-major
    for s, im in enumerate(images):
        out["image%d" % s] = np.asarray(im)
    return out


if __name__ == "__main__":
    blob = {}
    for name in CASES:
        for k, v in run_case(name).items():
            blob[name + "/" + k] = v
        print(name, "polygons per env:", blob[name + "/pair_n_polygons"].sum(axis=1).tolist())
    np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "steps.npz"), **blob)
