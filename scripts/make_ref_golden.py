"""Generates tests/golden/ref_bvh_vectors.npz with the REFERENCE's own compiled ray caster (oracle/_ref, built from
/root/reference/mujoco_contact_surface_sensors/src/bvh.cpp by oracle/ref_shim/Makefile).  Run in the container that
holds /root/reference:  python scripts/make_ref_golden.py

Contents (all made by reference code, none by the oracle's restatement of it):
* prim_*: known-answer vectors of IntersectTriangle (bvh.cpp:49-74) and IntersectAABB (bvh.h:157-176) incl. edge cases;
* per case c: triangle soups of the contact surfaces that touch the sensor (doubles, as Drake's tri_mesh_W() holds them;
  produced by the oracle's contact query: the Drake share, unpinned), a sample of the sensor's rays with the hits
  (t, u, v, (blas << 20) + triangle) the reference's BVH/TLAS returns for them — scalar and -DUSE_SSE builds —, and the
  full taxel image obtained with the reference's ray caster inside the flat-sensor loop.
The scenes are rebuilt from (presser, resolution, S, seed, env), so only poses-by-seed are stored implicitly.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mujoco_contact_surfaces_b200 import scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402
from parity_utils import make_oracle  # noqa: E402

CASES = [  # (name, factory kwargs, seed, env): the corners of the benchmark grid (benchmark_flat.cpp:282-373) + TLAS cases
    ("box_r25_s4", dict(presser="box", resolution=0.025, S=4), 7, 0),
    ("box_r25_s32", dict(presser="box", resolution=0.025, S=32), 7, 1),
    ("plate_r25_s8", dict(presser="plate", resolution=0.025, S=8), 7, 0),
    ("spot_r25_s32", dict(presser="spot", resolution=0.025, S=32), 7, 1),
    ("plate_r2p5_s4", dict(presser="plate", resolution=0.0025, S=4), 7, 0),
    ("spot_r2p5_s4", dict(presser="spot", resolution=0.0025, S=4), 7, 1),
    ("soft_tip_r25_s8", dict(presser="soft_tip", resolution=0.025, S=8), 7, 0),
    ("multi_r25_s8", dict(presser="multi", resolution=0.025, S=8), 7, 3),
]
MAX_RAYS = 6000


def make_scene(presser, resolution, S):
    if presser == "multi":
        return scenes.myrmex_multi(sampling_resolution=S, resolution=resolution)
    return scenes.myrmex(presser, S, resolution=resolution)


def soups(o, scene, sensor_geom):
    """Triangle soups of the surfaces that touch the sensor geom, in pair order (= BLAS order)."""
    n_tri, verts = [], []
    for p in range(len(scene.pairs)):
        r = o.pair_result(p)
        if r["has_surface"] and sensor_geom in (r["gM"], r["gN"]):
            t = o.pair_triangles(p)  # [n][12]: 9 vertex coordinates + 3 pressures
            n_tri.append(len(t))
            verts.append(t[:, :9].reshape(-1, 3, 3))
    return np.array(n_tri, np.int32), np.concatenate(verts)


def primitives(rng, R):
    import ctypes as C
    n = 4000
    O3 = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    D3 = rng.normal(size=(n, 3)).astype(np.float32)
    D3 /= np.linalg.norm(D3, axis=1, keepdims=True).astype(np.float32)
    tri = rng.uniform(-1, 1, (n, 3, 3)).astype(np.float32)
    t_in = np.where(rng.uniform(size=n) < 0.5, np.float32(1e30), rng.uniform(0, 2, n).astype(np.float32)).astype(np.float32)
    # edge cases: rays through a vertex / along an edge / parallel to the plane / axis-aligned directions
    for i in range(0, 400, 4):
        D3[i] = [0, 0, -1]
        O3[i] = [tri[i, 0, 0], tri[i, 0, 1], 2.0]  # through vertex 0
        O3[i + 1] = np.float32(0.5) * (tri[i + 1, 0] + tri[i + 1, 1]) - D3[i + 1]  # through the edge v0-v1
        e = tri[i + 2, 1] - tri[i + 2, 0]
        D3[i + 2] = e / np.linalg.norm(e)  # in the triangle's plane
        D3[i + 3] = np.eye(3, dtype=np.float32)[i % 3]
    tuv, hit = np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    for i in range(n):
        h = C.c_int(0)
        R.ref_intersect_triangle(fp(O3[i]), fp(D3[i]), fp(tri[i, 0]), fp(tri[i, 1]), fp(tri[i, 2]), C.c_float(t_in[i]),
                                 fp(tuv[i]), C.byref(h))
        hit[i] = h.value
    lo = rng.uniform(-1, 0.5, (n, 3)).astype(np.float32)
    hi = (lo + rng.uniform(0, 1, (n, 3))).astype(np.float32)
    hi[::7] = lo[::7]  # degenerate (flat) boxes
    Db = D3.copy()
    Db[::5, 0] = 0  # zero direction components: rD = inf
    Db[::11, 1] = -0.0
    tmin = np.array([R.ref_intersect_aabb(fp(O3[i]), fp(Db[i]), C.c_float(t_in[i]), fp(lo[i]), fp(hi[i])) for i in range(n)],
                    np.float32)
    return dict(prim_O=O3, prim_D=D3, prim_tri=tri, prim_t_in=t_in, prim_tuv=tuv, prim_hit=hit, prim_box_D=Db,
                prim_box_lo=lo, prim_box_hi=hi, prim_box_tmin=tmin)


def main():
    assert O.build_ref(), "needs /root/reference (run in the build container)"
    out = {}
    rng = np.random.Generator(np.random.PCG64(20261017))
    with np.errstate(divide="ignore", invalid="ignore"):
        out.update(primitives(rng, O.ref_lib(False)))
    names = []
    for name, kw, seed, env in CASES:
        sc = make_scene(kw["presser"], kw["resolution"], kw["S"])
        o = make_oracle(sc)
        xpos, xmat, vel = sc.poses(env + 1, seed=seed)
        o.step(xpos[env], xmat[env], vel[env])
        sensor_geom = sc.sensors[0]["geom"]
        n_tri, verts = soups(o, sc, sensor_geom)
        O.use_reference_caster(False)
        img, rays, tuv, hid = o.sensor_image_trace(0, kw["S"], 2)
        O.use_reference_caster(True)
        img_sse = o.sensor_image(0, 2)
        # every hit ray + a random sample of the misses, capped
        hits = np.nonzero(tuv[:, 0] < 1e30)[0]
        miss = np.nonzero(tuv[:, 0] >= 1e30)[0]
        if len(hits) > MAX_RAYS * 3 // 4:
            hits = np.sort(rng.choice(hits, MAX_RAYS * 3 // 4, replace=False))
        miss = np.sort(rng.choice(miss, min(len(miss), MAX_RAYS - len(hits)), replace=False))
        sel = np.sort(np.concatenate([hits, miss]))
        # the stand-alone entry must agree with the in-loop one
        tuv_s, id_s = O.ref_cast_rays(n_tri, verts, rays[sel, :3], rays[sel, 3:], sse=False)
        assert np.array_equal(tuv_s.view(np.uint32), tuv[sel].view(np.uint32)) and np.array_equal(id_s, hid[sel])
        tuv_e, id_e = O.ref_cast_rays(n_tri, verts, rays[sel, :3], rays[sel, 3:], sse=True)
        out.update({name + "_n_tri": n_tri, name + "_verts": verts, name + "_ray_index": sel.astype(np.int32),
                    name + "_rays": rays[sel], name + "_tuv": tuv_s, name + "_id": id_s, name + "_tuv_sse": tuv_e,
                    name + "_id_sse": id_e, name + "_image": img, name + "_image_sse": img_sse,
                    name + "_meta": np.array([kw["resolution"], kw["S"], seed, env], np.float64)})
        names.append(name + ":" + kw["presser"])
        print("%-16s surfaces %s rays %d (hits %d) image peak %.4g  sse image identical: %s, sse hits identical: %s" % (
            name, n_tri.tolist(), len(sel), len(hits), img.max(), np.array_equal(img, img_sse),
            np.array_equal(id_s, id_e)))
    out["cases"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "ref_bvh_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
