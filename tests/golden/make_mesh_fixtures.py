"""Generates tests/golden/meshes.npz from the reference's example mesh assets (run in the build container,
where /root/reference is mounted; the GPU box never reads /root/reference).

MuJoCo would hand the plugin mjModel.mesh_vert (float32) / mesh_face (int32)
(mujoco_contact_surfaces_plugin.cpp:745-763); MuJoCo is absent here, so the fixtures hold the raw asset
vertices with exact duplicates merged (what MuJoCo's STL loader does), without MuJoCo's re-centring.
"""
import struct
import sys

import numpy as np

REF = "/root/reference/mujoco_contact_surface_sensors/assets/meshes/"


def read_stl(path):
    raw = open(path, "rb").read()
    n = struct.unpack("<I", raw[80:84])[0]
    tris = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
    return tris["v"].reshape(-1, 3).copy()


def dedup(soup):
    uniq, inv = np.unique(soup, axis=0, return_inverse=True)
    # keep first-appearance order so the fixture does not depend on numpy's sort
    first = np.full(len(uniq), len(soup), dtype=np.int64)
    np.minimum.at(first, inv.reshape(-1), np.arange(len(soup)))
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(len(uniq))
    return uniq[order].astype(np.float32), rank[inv.reshape(-1)].reshape(-1, 3).astype(np.int32)


def read_obj(path):
    v, f = [], []
    for line in open(path):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            v.append([float(x) for x in p[1:4]])
        elif p[0] == "f":
            idx = [int(t.split("/")[0]) - 1 for t in p[1:]]
            for k in range(1, len(idx) - 1):
                f.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(v, dtype=np.float32), np.asarray(f, dtype=np.int32)


def signed_volume(v, f):
    a, b, c = v[f[:, 0]].astype(np.float64), v[f[:, 1]].astype(np.float64), v[f[:, 2]].astype(np.float64)
    return np.einsum("ij,ij->i", np.cross(a, b), c).sum() / 6


if __name__ == "__main__":
    out = {}
    pv, pf = dedup(read_stl(REF + "plate_8in_col.stl"))
    tv, tf = dedup(read_stl(REF + "ubi_tip_collision.stl"))
    sv, sf = read_obj(REF + "spot/spot_triangulated.obj")
    for name, (v, f) in dict(plate=(pv, pf), ubi_tip=(tv, tf), spot=(sv, sf)).items():
        print(name, v.shape, f.shape, "bbox", v.min(0), v.max(0), "volume", signed_volume(v, f))
        out[name + "_vert"], out[name + "_face"] = v, f
    np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/meshes.npz", **out)
