"""Pins the oracle's restatement of the tactile ray caster to REFERENCE-HELD code.

The reference's own float32 BVH / TLAS / Moeller-Trumbore / slab test (mujoco_contact_surface_sensors/src/bvh.cpp:49-476,
include/.../bvh.h:69-281, float3.h) compiles unmodified against the container-only shim headers of oracle/ref_shim into
oracle/_ref/ (scalar and -DUSE_SSE builds).  Two layers:

* tests/golden/ref_bvh_vectors.npz — inputs and outputs of that compiled reference code, made by
  scripts/make_ref_golden.py: always checked, on any machine;
* the live libraries under oracle/_ref/ — checked when present (here, and on the GPU box, where the prebuilt .so travels).

Bars: hits, misses, (t, u, v) and triangle ids BIT-EXACT against the scalar build; against the SSE build the taxel images
are bit-identical and rays may differ only by WHICH of two triangles sharing the hit edge is reported (same t): the
reference's own two builds differ from each other in exactly that way.
"""
import os

import numpy as np
import pytest

from mujoco_contact_surfaces_b200 import scenes
from oracle import oracle as O
from parity_utils import make_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_bvh_vectors.npz")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLDEN) as z:  # NpzFile decompresses on every access: read it once
        return {k: z[k] for k in z.files}


def _cases():
    return [c.split(":") for c in np.load(GOLDEN)["cases"]]


def _scene(presser, resolution, S):
    if presser == "multi":
        return scenes.myrmex_multi(sampling_resolution=S, resolution=resolution)
    return scenes.myrmex(presser, S, resolution=resolution)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_primitives_match_the_reference_vectors(gold):
    """IntersectTriangle (bvh.cpp:49-74) and IntersectAABB (bvh.h:157-176), 4000 cases each incl. rays through vertices
    and edges, rays in the triangle's plane, flat boxes, zero direction components (rD = +-inf)."""
    import ctypes as C
    L = O.lib()
    fp = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    n = len(gold["prim_O"])
    tuv, hit, tmin = np.zeros((n, 3), np.float32), np.zeros(n, np.int32), np.zeros(n, np.float32)
    with np.errstate(all="ignore"):
        for i in range(n):
            h = C.c_int(0)
            L.orc_intersect_triangle(fp(gold["prim_O"][i]), fp(gold["prim_D"][i]), fp(gold["prim_tri"][i, 0]),
                                     fp(gold["prim_tri"][i, 1]), fp(gold["prim_tri"][i, 2]), C.c_float(gold["prim_t_in"][i]),
                                     fp(tuv[i]), C.byref(h))
            hit[i] = h.value
            tmin[i] = L.orc_intersect_aabb(fp(gold["prim_O"][i]), fp(gold["prim_box_D"][i]), C.c_float(gold["prim_t_in"][i]),
                                           fp(gold["prim_box_lo"][i]), fp(gold["prim_box_hi"][i]))
    assert 100 < gold["prim_hit"].sum() < n and np.array_equal(hit, gold["prim_hit"])
    assert np.array_equal(_bits(tuv), _bits(gold["prim_tuv"]))
    assert np.array_equal(_bits(tmin), _bits(gold["prim_box_tmin"]))
    assert 100 < (gold["prim_box_tmin"] < 1e30).sum() < n


@pytest.mark.parametrize("name,presser", _cases())
def test_oracle_ray_caster_matches_the_reference_vectors(gold, name, presser):
    """BLAS build (binned SAH), TLAS build (agglomerative) and both traversals: same soups + same rays in, bit-identical
    nearest hits out."""
    n_tri, verts, rays = gold[name + "_n_tri"], gold[name + "_verts"], gold[name + "_rays"]
    tuv, hid = O.oracle_cast_rays(n_tri, verts, rays[:, :3], rays[:, 3:])
    assert (gold[name + "_tuv"][:, 0] < 1e30).sum() > 30
    assert np.array_equal(_bits(tuv), _bits(gold[name + "_tuv"])), "t/u/v differ from the reference's scalar build"
    assert np.array_equal(hid, gold[name + "_id"]), "triangle ids differ from the reference's scalar build"
    # the reference's -DUSE_SSE build (its CMake default): same hit/miss pattern, same t; ids/(u,v) may differ only on ties
    t_sse, id_sse = gold[name + "_tuv_sse"], gold[name + "_id_sse"]
    assert np.array_equal(_bits(tuv[:, 0]), _bits(t_sse[:, 0]))
    differ = np.nonzero(hid != id_sse)[0]
    assert len(differ) <= 0.02 * len(hid)
    for i in differ:  # a tie: the ray runs through an edge shared by the two triangles
        assert min(tuv[i, 1], tuv[i, 2], 1 - tuv[i, 1] - tuv[i, 2]) <= 1e-6


@pytest.mark.parametrize("name,presser", _cases())
def test_oracle_flat_sensor_image_matches_the_reference_caster(gold, name, presser):
    """Whole flat-sensor loop: the oracle (its own BVH, and the linear scan) against the image made with the reference's
    compiled ray caster; also guards the stored soups against drift of the oracle's contact query."""
    res, S, seed, env = gold[name + "_meta"]
    S, seed, env = int(S), int(seed), int(env)
    sc = _scene(presser, float(res), S)
    o = make_oracle(sc)
    xpos, xmat, vel = sc.poses(env + 1, seed=seed)
    o.step(xpos[env], xmat[env], vel[env])
    img, rays, tuv, hid = o.sensor_image_trace(0, S, 1)
    assert np.array_equal(_bits(img), _bits(gold[name + "_image"]))
    assert np.array_equal(_bits(img), _bits(gold[name + "_image_sse"]))
    sel = gold[name + "_ray_index"]
    assert np.array_equal(_bits(rays[sel]), _bits(gold[name + "_rays"]))
    assert np.array_equal(_bits(tuv[sel]), _bits(gold[name + "_tuv"])) and np.array_equal(hid[sel], gold[name + "_id"])
    if S <= 8:  # the linear scan may report the other triangle of a tie; the image must not care beyond rounding
        lin = o.sensor_image(0, 0)
        assert np.allclose(lin, img, rtol=1e-6, atol=0)


needs_ref = pytest.mark.skipif(not (O.ref_available(False) and O.ref_available(True)),
                               reason="oracle/_ref not built (needs /root/reference; prebuilt .so files travel to the GPU box)")


@needs_ref
@pytest.mark.parametrize("sse", [False, True])
@pytest.mark.parametrize("presser,resolution,S", [("box", 0.025, 4), ("box", 0.025, 20), ("plate", 0.025, 8), ("spot", 0.025, 16),
                                                   ("plate", 0.0025, 4), ("soft_tip", 0.025, 8), ("multi", 0.025, 8)])
def test_live_reference_ray_caster(presser, resolution, S, sse):
    """oracle/_ref live: fresh seeds, three environments per case."""
    O.use_reference_caster(sse)
    sc = _scene(presser, resolution, S)
    o = make_oracle(sc)
    xpos, xmat, vel = sc.poses(3, seed=2026)
    n_hits = 0
    for e in range(3):
        o.step(xpos[e], xmat[e], vel[e])
        img1, _, tuv1, id1 = o.sensor_image_trace(0, S, 1)
        img2, _, tuv2, id2 = o.sensor_image_trace(0, S, 2)
        assert np.array_equal(_bits(img1), _bits(img2))
        assert np.array_equal(_bits(tuv1[:, 0]), _bits(tuv2[:, 0]))
        if not sse:
            assert np.array_equal(_bits(tuv1), _bits(tuv2)) and np.array_equal(id1, id2)
        else:
            d = np.nonzero(id1 != id2)[0]
            assert len(d) <= 0.02 * max(1, (tuv1[:, 0] < 1e30).sum())
            assert all(min(tuv1[i, 1], tuv1[i, 2], 1 - tuv1[i, 1] - tuv1[i, 2]) <= 1e-6 for i in d)
        n_hits += int((tuv1[:, 0] < 1e30).sum())
    assert n_hits > 20


@needs_ref
@pytest.mark.parametrize("with_normals", [True, False])
def test_live_reference_ray_caster_under_the_curved_sensor(with_normals):
    """CurvedSensor::internal_update (curved_sensor.cpp:388-481) casts its rays through the same BVH/TLAS."""
    O.use_reference_caster(False)
    sc = scenes.fingertip(n_samples=1500, with_normals=with_normals)
    o = make_oracle(sc)
    xpos, xmat, vel = sc.poses(4, seed=11)
    seen = 0
    for e in range(4):
        o.step(xpos[e], xmat[e], vel[e])
        own, ref = o.curved_values(0, 1), o.curved_values(0, 2)
        assert np.array_equal(_bits(own), _bits(ref))
        seen += int((own > 0).sum())
    assert seen > 0
