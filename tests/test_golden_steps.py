"""Committed step fixtures (tests/golden/steps.npz, written by tests/golden/make_step_fixtures.py).

CPU: the oracle and the scene generators still reproduce the file (a regression pin on the checker itself).
GPU: the CUDA path, through the C ABI, against the file on the file's own inputs — candidate sets and vertex counts
exactly, forces/torques/centroids within 1e-8, taxel images within 1e-6 (the north-star bars)."""
import importlib.util
import os

import numpy as np
import pytest

from parity_utils import FORCE_RTOL, compare_images, make_engine, make_oracle, oracle_env, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_step_fixtures", os.path.join(HERE, "golden", "make_step_fixtures.py"))
fixtures = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fixtures)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "steps.npz"))


def _emitted(golden, name, e, p, n_pairs):
    off = golden[name + "/emitted_off"]
    i = e * n_pairs + p
    return set(map(tuple, golden[name + "/emitted"][off[i]:off[i + 1]].tolist()))


@pytest.mark.parametrize("name", sorted(fixtures.CASES))
def test_oracle_and_scene_generators_reproduce_the_fixtures(golden, name):
    factory, n_envs, seed, sensors = fixtures.CASES[name]
    scene = factory()
    xpos, xmat, vel = scene.poses(n_envs, seed)
    for k, v in (("xpos", xpos), ("xmat", xmat), ("vel", vel)):
        assert np.array_equal(v, golden[name + "/" + k]), "scene generator changed: " + k
    orc = make_oracle(scene)
    n_pairs = len(scene.pairs)
    for e in range(n_envs):
        pairs, imgs = oracle_env(orc, scene, xpos[e], xmat[e], vel[e], sensors=sensors)
        for p in range(n_pairs):
            assert pairs[p]["emitted"] == _emitted(golden, name, e, p, n_pairs)
            for f in ("n_polygons", "n_faces", "n_points"):
                assert pairs[p][f] == int(golden[name + "/pair_" + f][e, p])
            for f in ("F", "tau", "centroid", "area"):  # 1e-12: another libm / compiler may move the last bits
                assert rel_err(pairs[p][f], golden[name + "/pair_" + f][e, p], floor=1e-300) < 1e-12, (f, e, p)
        for s, img in enumerate(imgs):
            np.testing.assert_allclose(img, golden[name + "/image%d" % s][e], rtol=1e-6, atol=0)
        for s in range(len(getattr(scene, "taxel_sensors", []))):
            ref = golden[name + "/taxel%d" % s][e]
            assert ref.max() > 0
            np.testing.assert_allclose(orc.taxel_values(s), ref, rtol=1e-6, atol=1e-6 * ref.max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(fixtures.CASES))
def test_golden_steps(hcs_lib, golden, name):
    factory, n_envs, seed, sensors = fixtures.CASES[name]
    scene = factory()
    eng = make_engine(scene, n_envs)
    n_taxel = len(getattr(scene, "taxel_sensors", []))
    eng.step(golden[name + "/xpos"], golden[name + "/xmat"], golden[name + "/vel"], with_sensors=sensors or n_taxel > 0)
    res, wrench = eng.pair_results(), eng.geom_wrenches()
    n_pairs = len(scene.pairs)
    n_poly = 0
    for e in range(n_envs):
        for p in range(n_pairs):
            g = res[e][p]
            assert set(map(tuple, eng.emitted(e, p).tolist())) == _emitted(golden, name, e, p, n_pairs)
            for f in ("n_polygons", "n_faces", "n_points"):
                assert int(g[f]) == int(golden[name + "/pair_" + f][e, p]), (f, e, p)
            n_poly += int(g["n_polygons"])
            if not golden[name + "/pair_has_surface"][e, p]:
                assert np.all(g["F"] == 0) and np.all(g["tau"] == 0)
                continue
            F, tau = golden[name + "/pair_F"][e, p], golden[name + "/pair_tau"][e, p]
            fscale = np.linalg.norm(F)
            if fscale > 0:
                assert rel_err(g["F"], F) < FORCE_RTOL
                assert np.linalg.norm(g["tau"] - tau) / max(np.linalg.norm(tau), 0.1 * fscale) < FORCE_RTOL
            assert abs(g["area"] - golden[name + "/pair_area"][e, p]) / golden[name + "/pair_area"][e, p] < FORCE_RTOL
            assert rel_err(g["centroid"], golden[name + "/pair_centroid"][e, p], floor=1e-3) < FORCE_RTOL
        for gi in range(scene.n_geoms):
            ref = golden[name + "/geom_wrench"][e, gi]
            assert np.linalg.norm(wrench[e, gi] - ref) <= 1e-8 * max(np.linalg.norm(ref), 1e-9)
        if sensors:
            for s in range(len(scene.sensors)):
                err, nbad = compare_images(eng.sensor_image(s)[e], golden[name + "/image%d" % s][e])
                assert nbad == 0, "taxel image: %d taxels off (max rel err %.3e)" % (nbad, err)
        for s in range(n_taxel):
            err, nbad = compare_images(eng.taxel_values(s)[e], golden[name + "/taxel%d" % s][e])
            assert nbad == 0, "taxel sensor: %d taxels off (max rel err %.3e)" % (nbad, err)
    assert n_poly > 0
    eng.close()
