"""ctypes binding of the CPU oracle (oracle/liboracle_hcs.so).

TEST INFRASTRUCTURE ONLY — parity unpinned (see oracle/oracle.hpp).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
KIND_RIGID, KIND_SOFT, KIND_PLANE = 0, 1, 2


def build(force=False):
    so = os.path.join(_HERE, "liboracle_hcs.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_scene_create.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_scene_destroy", "orc_add_geom", "orc_add_raw_soft", "orc_add_raw_rigid", "orc_geom_info",
                     "orc_geom_mesh", "orc_set_pairs", "orc_add_flat_sensor", "orc_sensor_dims", "orc_step",
                     "orc_pair_result", "orc_pair_emitted", "orc_pair_faces", "orc_pair_triangles", "orc_geom_wrench",
                     "orc_pair_face_vertices",
                     "orc_sensor_image", "orc_bench", "orc_add_curved_sensor", "orc_curved_values", "orc_curved_info",
                     "orc_add_taxel_sensor", "orc_taxel_values", "orc_sensor_image_trace", "orc_set_external_caster",
                     "orc_cast_rays"):
            getattr(L, name).restype = C.c_int
        L.orc_set_external_caster.argtypes = [C.c_void_p] * 3
        L.orc_intersect_aabb.restype = C.c_float
        L.orc_intersect_aabb.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float)]
        L.orc_intersect_triangle.argtypes = [C.POINTER(C.c_float)] * 5 + [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.orc_intersect_triangle.restype = None
        _LIB = L
    return _LIB


REF_SOURCE = "/root/reference/mujoco_contact_surface_sensors"
_REF = {}


def build_ref():
    """Compiles the REFERENCE's own ray caster (bvh.cpp, unmodified, where it lies under /root/reference) against the
    shim headers of oracle/ref_shim into oracle/_ref/ — only where the reference is present (this container).  On the
    GPU box the prebuilt libraries that travelled with the snapshot are used.  Returns True when oracle/_ref exists."""
    if os.path.exists(os.path.join(REF_SOURCE, "src", "bvh.cpp")):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "ref_shim"), "-s"])
    return ref_available()


def ref_available(sse=False):
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_bvh_sse.so" if sse else "libref_bvh.so"))


def ref_lib(sse=False):
    """oracle/_ref/libref_bvh.so (scalar code path) or libref_bvh_sse.so (-DUSE_SSE, the reference's CMake default)."""
    if sse not in _REF:
        R = C.CDLL(os.path.join(_HERE, "_ref", "libref_bvh_sse.so" if sse else "libref_bvh.so"))
        R.ref_tlas_create.restype = C.c_void_p
        R.ref_tlas_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        R.ref_tlas_cast.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(C.c_uint32)]
        R.ref_tlas_destroy.argtypes = [C.c_void_p]
        R.ref_evaluate.restype = C.c_double
        R.ref_intersect_aabb.restype = C.c_float
        R.ref_intersect_aabb.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float)]
        R.ref_intersect_triangle.argtypes = [C.POINTER(C.c_float)] * 5 + [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        assert R.ref_use_sse() == int(sse)
        _REF[sse] = R
    return _REF[sse]


def use_reference_caster(sse=False):
    """Routes caster mode 2 of the oracle's sensors (sensor_image(use_bvh=2), curved_values(use_bvh=2)) through the
    reference's compiled BVH/TLAS."""
    R, L = ref_lib(sse), lib()
    L.orc_set_external_caster(C.cast(R.ref_tlas_create, C.c_void_p), C.cast(R.ref_tlas_cast, C.c_void_p),
                              C.cast(R.ref_tlas_destroy, C.c_void_p))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ref_cast_rays(n_tri, verts, O, D, sse=False):
    """The reference's ray caster on triangle soups: n_tri[s] triangles per surface, verts [sum n_tri][3][3] doubles."""
    R = ref_lib(sse)
    n_tri, verts, O, D = np.ascontiguousarray(n_tri, dtype=np.int32), _d(verts), _f(O), _f(D)
    n = len(O)
    tuv, hid = np.empty((n, 3), np.float32), np.empty(n, np.uint32)
    h = C.c_void_p(R.ref_tlas_create(len(n_tri), _p(n_tri, C.c_int), _p(verts, C.c_double), None))
    R.ref_tlas_cast(h, n, _p(O, C.c_float), _p(D, C.c_float), _p(tuv, C.c_float), _p(hid, C.c_uint32))
    R.ref_tlas_destroy(h)
    return tuv, hid


def oracle_cast_rays(n_tri, verts, O, D):
    """The oracle's restatement of the same ray caster, same arguments."""
    L = lib()
    n_tri, verts, O, D = np.ascontiguousarray(n_tri, dtype=np.int32), _d(verts), _f(O), _f(D)
    n = len(O)
    tuv, hid = np.empty((n, 3), np.float32), np.empty(n, np.uint32)
    L.orc_cast_rays(len(n_tri), _p(n_tri, C.c_int), _p(verts, C.c_double), n, _p(O, C.c_float), _p(D, C.c_float),
                    _p(tuv, C.c_float), _p(hid, C.c_uint32))
    return tuv, hid


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleScene:
    """One mjModel worth of hydroelastic configuration + one env of per-step state."""

    def __init__(self, triangle_representation=False, apply_forces=True):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_scene_create(int(triangle_representation), int(apply_forces)))
        self.n_geoms = 0
        self.n_pairs = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def add_geom(self, mj_type, size, props, mesh_vert=None, mesh_face=None):
        size = _d(np.resize(np.asarray(size, dtype=np.float64), 3))
        props = _d(props)
        mv = np.ascontiguousarray(mesh_vert, dtype=np.float32) if mesh_vert is not None else None
        mf = np.ascontiguousarray(mesh_face, dtype=np.int32) if mesh_face is not None else None
        r = self.L.orc_add_geom(self.h, int(mj_type), _p(size, C.c_double), _p(mv, C.c_float),
                                0 if mv is None else len(mv), _p(mf, C.c_int), 0 if mf is None else len(mf),
                                _p(props, C.c_double))
        if r < 0:
            raise ValueError(self.L.orc_last_error().decode())
        self.n_geoms = r + 1
        return r

    def add_raw_soft(self, verts, tets, pressure, props):
        verts, pressure, props = _d(verts), _d(pressure), _d(props)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        r = self.L.orc_add_raw_soft(self.h, _p(verts, C.c_double), len(verts), _p(tets, C.c_int), len(tets),
                                    _p(pressure, C.c_double), _p(props, C.c_double))
        self.n_geoms = r + 1
        return r

    def add_raw_rigid(self, verts, tris, props):
        verts, props = _d(verts), _d(props)
        tris = np.ascontiguousarray(tris, dtype=np.int32)
        r = self.L.orc_add_raw_rigid(self.h, _p(verts, C.c_double), len(verts), _p(tris, C.c_int), len(tris),
                                     _p(props, C.c_double))
        self.n_geoms = r + 1
        return r

    def geom_mesh(self, gi):
        info = np.zeros(3, dtype=np.int32)
        assert self.L.orc_geom_info(self.h, gi, _p(info, C.c_int)) == 0
        kind, nv, ne = (int(x) for x in info)
        if kind == KIND_PLANE:
            return dict(kind=kind)
        per = 4 if kind == KIND_SOFT else 3
        verts = np.zeros((nv, 3))
        elems = np.zeros((ne, per), dtype=np.int32)
        grad = np.zeros((ne, 3))
        pressure = np.zeros(nv) if kind == KIND_SOFT else None
        e0 = np.zeros(ne) if kind == KIND_SOFT else None
        self.L.orc_geom_mesh(self.h, gi, _p(verts, C.c_double), _p(elems, C.c_int), _p(pressure, C.c_double),
                             _p(grad, C.c_double), _p(e0, C.c_double))
        out = dict(kind=kind, verts=verts, elems=elems)
        if kind == KIND_SOFT:
            out.update(pressure=pressure, grad=grad, e0=e0)
        else:
            out.update(normal=grad)
        return out

    def set_pairs(self, pairs):
        pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        g1, g2 = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
        assert self.L.orc_set_pairs(self.h, _p(g1, C.c_int), _p(g2, C.c_int), len(pairs)) == 0
        self.n_pairs = len(pairs)

    def add_flat_sensor(self, geom, geom_size, resolution, sampling_resolution, window=0, sigma=-1.0):
        gs = _d(geom_size)
        r = self.L.orc_add_flat_sensor(self.h, int(geom), _p(gs, C.c_double), C.c_double(resolution),
                                       int(sampling_resolution), int(window), C.c_float(sigma))
        return r

    def add_curved_sensor(self, geom, taxel_pos, taxel_nrm, sample_pos, sample_nrm, include_margin):
        """CurvedSensor::load with caller-supplied surface samples; returns the sensor index."""
        tp, sp, sn = _d(taxel_pos).reshape(-1, 3), _d(sample_pos).reshape(-1, 3), _d(sample_nrm).reshape(-1, 3)
        tn = None if taxel_nrm is None else _d(taxel_nrm).reshape(-1, 3)
        s = self.L.orc_add_curved_sensor(self.h, int(geom), len(tp), _p(tp, C.c_double),
                                         None if tn is None else _p(tn, C.c_double), len(sp), _p(sp, C.c_double),
                                         _p(sn, C.c_double), C.c_double(include_margin))
        if not hasattr(self, "curved_taxels"):
            self.curved_taxels = []
        self.curved_taxels.append(len(tp))
        return s

    def curved_values(self, sensor, use_bvh=True):
        out = np.zeros(self.curved_taxels[sensor], dtype=np.float32)
        self.L.orc_curved_values(self.h, int(sensor), _p(out, C.c_float), int(use_bvh))
        return out

    def curved_info(self, sensor):
        d = np.zeros(2, dtype=np.int32)
        self.L.orc_curved_info(self.h, int(sensor), _p(d, C.c_int))
        return int(d[0]), int(d[1])  # close sample points, (taxel, sample) assignments

    def add_taxel_sensor(self, geom, taxel_pos, include_margin, sample_resolution, method="squared", visualize=False,
                         sample_method="default"):
        """TaxelSensor::load; method: closest | weighted | mean | squared; sample_method: default | area_importance."""
        tp = _d(taxel_pos).reshape(-1, 3)
        code = {"closest": 0, "weighted": 1, "mean": 2, "squared": 3}[method]
        sm = {"default": 0, "area_importance": 1}[sample_method]
        s = self.L.orc_add_taxel_sensor(self.h, int(geom), len(tp), _p(tp, C.c_double), C.c_double(include_margin),
                                        C.c_double(sample_resolution), code, int(visualize), sm)
        if not hasattr(self, "taxel_counts"):
            self.taxel_counts = []
        self.taxel_counts.append(len(tp))
        return s

    def taxel_values(self, sensor, previous=None):
        """One update of the sensor; `previous` = the message of the last update (taxels without a sample in range
        keep it), zeros when omitted."""
        out = np.zeros(self.taxel_counts[sensor], dtype=np.float32) if previous is None else \
            np.array(previous, dtype=np.float32).copy()
        self.L.orc_taxel_values(self.h, int(sensor), _p(out, C.c_float))
        return out

    def sensor_dims(self, sensor):
        d = np.zeros(2, dtype=np.int32)
        self.L.orc_sensor_dims(self.h, sensor, _p(d, C.c_int))
        return int(d[0]), int(d[1])

    def step(self, xpos, xmat, vel=None, use_bvh=True):
        xpos, xmat = _d(xpos).reshape(-1), _d(xmat).reshape(-1)
        vel = np.zeros(6 * self.n_geoms) if vel is None else _d(vel).reshape(-1)
        assert xpos.size == 3 * self.n_geoms and xmat.size == 9 * self.n_geoms and vel.size == 6 * self.n_geoms
        if self.L.orc_step(self.h, _p(xpos, C.c_double), _p(xmat, C.c_double), _p(vel, C.c_double), int(use_bvh)):
            raise RuntimeError(self.L.orc_last_error().decode())

    def pair_result(self, pair):
        o = np.zeros(17)
        self.L.orc_pair_result(self.h, pair, _p(o, C.c_double))
        return dict(has_surface=bool(o[0]), gM=int(o[1]), gN=int(o[2]), n_faces=int(o[3]), n_polygons=int(o[4]),
                    n_points=int(o[5]), n_candidates=int(o[6]), F=o[7:10].copy(), tau=o[10:13].copy(),
                    centroid=o[13:16].copy(), area=float(o[16]))

    def pair_emitted(self, pair):
        n = self.L.orc_pair_emitted(self.h, pair, None, 0)
        buf = np.zeros((max(n, 1), 3), dtype=np.int32)
        self.L.orc_pair_emitted(self.h, pair, _p(buf, C.c_int), n)
        return buf[:n]

    def pair_faces(self, pair):
        n = self.L.orc_pair_faces(self.h, pair, None, 0)
        buf = np.zeros((max(n, 1), 13))
        self.L.orc_pair_faces(self.h, pair, _p(buf, C.c_double), n)
        return buf[:n]

    def pair_face_vertices(self, pair):
        """per PointCollision: (vertex count, (8, 3) world vertices of its face)"""
        n = self.L.orc_pair_face_vertices(self.h, pair, None, 0)
        buf = np.zeros((max(n, 1), 25))
        self.L.orc_pair_face_vertices(self.h, pair, _p(buf, C.c_double), n)
        return buf[:n, 0].astype(int), buf[:n, 1:].reshape(-1, 8, 3)

    def pair_triangles(self, pair):
        n = self.L.orc_pair_triangles(self.h, pair, None, 0)
        buf = np.zeros((max(n, 1), 12))
        self.L.orc_pair_triangles(self.h, pair, _p(buf, C.c_double), n)
        return buf[:n]

    def geom_wrench(self, geom):
        o = np.zeros(6)
        self.L.orc_geom_wrench(self.h, geom, _p(o, C.c_double))
        return o

    def sensor_image(self, sensor, use_bvh=True, parallel=False):
        """use_bvh: False/0 linear scan, True/1 the oracle's BVH restatement, 2 the reference's compiled ray caster
        (needs use_reference_caster() first)."""
        cx, cy = self.sensor_dims(sensor)
        img = np.zeros(cx * cy, dtype=np.float32)
        if self.L.orc_sensor_image(self.h, sensor, _p(img, C.c_float), int(use_bvh), int(parallel)):
            raise RuntimeError(self.L.orc_last_error().decode())
        return img

    def sensor_image_trace(self, sensor, S, use_bvh=True):
        """Image plus every ray (O, D) and its nearest hit (t, u, v, id) in (x, y, i, j) order."""
        cx, cy = self.sensor_dims(sensor)
        n = cx * cy * S * S
        img = np.zeros(cx * cy, dtype=np.float32)
        rays, tuv, hid = np.zeros((n, 6), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.uint32)
        if self.L.orc_sensor_image_trace(self.h, sensor, _p(img, C.c_float), int(use_bvh), _p(rays, C.c_float),
                                         _p(tuv, C.c_float), _p(hid, C.c_uint32)):
            raise RuntimeError(self.L.orc_last_error().decode())
        return img, rays, tuv, hid

    def bench(self, xpos, xmat, vel, use_bvh=True, with_sensors=False, threads=1):
        """Time n_env env steps on the host; returns (seconds, candidate pair-evals, checksum)."""
        xpos, xmat, vel = _d(xpos), _d(xmat), _d(vel)
        n_env = xpos.size // (3 * self.n_geoms)
        o = np.zeros(3)
        self.L.orc_bench(self.h, n_env, _p(xpos, C.c_double), _p(xmat, C.c_double), _p(vel, C.c_double),
                         int(use_bvh), int(with_sensors), int(threads), _p(o, C.c_double))
        return float(o[0]), int(o[1]), float(o[2])


def num_threads():
    return int(lib().orc_num_threads())
