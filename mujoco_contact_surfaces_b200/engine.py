"""ctypes binding of libhcs_b200.so (the C ABI declared in include/hcs.h).

The Python side is plumbing only (buffers, device selection, launching one process per GPU); every
number is produced by the hand-written CUDA kernels behind the C ABI.  There is no CPU fallback: loading
fails loudly when the library is missing, and hcs_create fails when no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HCS_LIB") or os.path.join(_HERE, "libhcs_b200.so")  # HCS_LIB: tuning variants

GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
REP_POLYGON, REP_TRIANGLE = 0, 1
WINDOW_NONE, WINDOW_GAUSS, WINDOW_TUKEY, WINDOW_SQUARE = range(4)
HCS_OK, HCS_E_INVALID, HCS_E_UNSUPPORTED, HCS_E_CUDA, HCS_E_CAPACITY, HCS_E_NOT_FINALIZED = 0, -1, -2, -3, -4, -5


class HcsConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("n_envs", C.c_int), ("representation", C.c_int),
                ("apply_contact_forces", C.c_int), ("max_candidates_per_slice", C.c_int), ("max_faces", C.c_int),
                ("max_tactile_triangles", C.c_int), ("max_triangles_per_taxel", C.c_int), ("stream", C.c_void_p),
                ("face_vertices", C.c_int)]


class HcsOutputs(C.Structure):
    """hcs_outputs: caller-owned host buffers of a pipelined step (include/hcs.h)."""
    _fields_ = [("geom_wrench", C.c_void_p), ("sensor_images", C.POINTER(C.c_void_p)),
                ("curved_values", C.POINTER(C.c_void_p)), ("taxel_values", C.POINTER(C.c_void_p)),
                ("pair_results", C.c_void_p)]


PAIR_RESULT_DTYPE = np.dtype([("F", "<f8", 3), ("tau", "<f8", 3), ("centroid", "<f8", 3), ("area", "<f8"),
                              ("gM", "<i4"), ("gN", "<i4"), ("n_polygons", "<i4"), ("n_faces", "<i4"),
                              ("n_points", "<i4"), ("n_candidates", "<i4"), ("n_clipped", "<i4"),
                              ("reserved", "<i4")], align=True)
FACE_DTYPE = np.dtype([("p", "<f8", 3), ("n", "<f8", 3), ("fn0", "<f8"), ("stiffness", "<f8"), ("damping", "<f8"),
                       ("f", "<f8", 3), ("env", "<i4"), ("pair", "<i4"), ("elemM", "<i4"), ("elemN", "<i4"),
                       ("nverts", "<i4"), ("face", "<i4")], align=True)
assert PAIR_RESULT_DTYPE.itemsize == 112 and FACE_DTYPE.itemsize == 120

# every symbol include/hcs.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "hcs_create", "hcs_destroy", "hcs_last_error", "hcs_add_geom", "hcs_add_soft_mesh", "hcs_add_rigid_mesh",
    "hcs_update_geom", "hcs_set_pairs", "hcs_add_flat_sensor", "hcs_sensor_dims", "hcs_finalize", "hcs_step",
    "hcs_step_device", "hcs_sync", "hcs_fetch_results", "hcs_n_geoms", "hcs_n_pairs", "hcs_get_pair_results",
    "hcs_get_geom_wrenches", "hcs_get_sensor_image", "hcs_device_pair_results", "hcs_device_geom_wrenches",
    "hcs_device_sensor_image", "hcs_get_faces", "hcs_get_emitted", "hcs_get_tactile_triangles", "hcs_geom_info",
    "hcs_get_mesh", "hcs_get_lbvh", "hcs_get_counters", "hcs_set_profiling", "hcs_get_stage_ms", "hcs_version",
    "hcs_add_curved_sensor", "hcs_curved_sensor_info", "hcs_get_curved_values", "hcs_device_curved_values",
    "hcs_add_taxel_sensor", "hcs_get_taxel_values", "hcs_device_taxel_values", "hcs_get_face_vertices",
    "hcs_update_flat_sensor", "hcs_step_async", "hcs_wait", "hcs_get_tactile_triangle_pairs", "hcs_set_env_sizes", "hcs_multi_set_env_sizes",
    "hcs_multi_create", "hcs_multi_destroy", "hcs_multi_last_error", "hcs_multi_n_blocks", "hcs_multi_block",
    "hcs_multi_add_geom", "hcs_multi_add_soft_mesh", "hcs_multi_add_rigid_mesh", "hcs_multi_update_geom",
    "hcs_multi_set_pairs", "hcs_multi_add_flat_sensor", "hcs_multi_finalize", "hcs_multi_step", "hcs_multi_step_async",
    "hcs_multi_wait", "hcs_multi_get_geom_wrenches", "hcs_multi_get_pair_results", "hcs_multi_get_sensor_image",
]

_LIB = None


class HcsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("hcs status %d: %s" % (status, message))
        self.status = status


def load_library():
    """dlopen the in-tree CUDA library; never falls back to anything else."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -m mujoco_contact_surfaces_b200.build` "
                              "(there is no CPU fallback for the contact path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.hcs_last_error.restype = C.c_char_p
        L.hcs_version.restype = C.c_char_p
        L.hcs_last_error.argtypes = [C.c_void_p]
        L.hcs_destroy.restype = None
        L.hcs_destroy.argtypes = [C.c_void_p]
        for name in ("hcs_device_pair_results", "hcs_device_geom_wrenches", "hcs_device_sensor_image",
                     "hcs_device_curved_values", "hcs_device_taxel_values"):
            getattr(L, name).restype = C.c_void_p
        L.hcs_device_taxel_values.argtypes = [C.c_void_p, C.c_int]
        L.hcs_add_taxel_sensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                           C.c_int, C.c_int]
        L.hcs_device_curved_values.argtypes = [C.c_void_p, C.c_int]
        L.hcs_add_curved_sensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_double]
        L.hcs_device_pair_results.argtypes = [C.c_void_p]
        L.hcs_device_geom_wrenches.argtypes = [C.c_void_p]
        L.hcs_device_sensor_image.argtypes = [C.c_void_p, C.c_int]
        L.hcs_step_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hcs_step_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hcs_wait.argtypes = [C.c_void_p, C.c_int64]
        L.hcs_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hcs_set_env_sizes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hcs_multi_set_env_sizes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hcs_multi_create.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.hcs_multi_destroy.argtypes = [C.c_void_p]
        L.hcs_multi_destroy.restype = None
        L.hcs_multi_last_error.argtypes = [C.c_void_p]
        L.hcs_multi_last_error.restype = C.c_char_p
        L.hcs_multi_n_blocks.argtypes = [C.c_void_p]
        L.hcs_multi_block.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        L.hcs_multi_add_geom.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_int,
                                         C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_double)]
        L.hcs_multi_set_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
        L.hcs_multi_add_flat_sensor.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_float]
        L.hcs_multi_update_geom.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.hcs_multi_finalize.argtypes = [C.c_void_p]
        L.hcs_multi_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hcs_multi_step_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hcs_multi_wait.argtypes = [C.c_void_p, C.c_int64]
        L.hcs_multi_get_geom_wrenches.argtypes = [C.c_void_p, C.c_void_p]
        L.hcs_multi_get_pair_results.argtypes = [C.c_void_p, C.c_void_p]
        L.hcs_multi_get_sensor_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hcs_add_flat_sensor.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_float]
        L.hcs_update_flat_sensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
        _LIB = L
    return _LIB


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class HydroelasticEngine:
    """Batched hydroelastic contact engine for `n_envs` independent environments on one GPU.

    Mirrors the configuration surface of MujocoContactSurfacesPlugin (parseMujocoCustomFields,
    mujoco_contact_surfaces_plugin.cpp:571-813): add_geom() once per `cs::<geom>` numeric in XML order,
    set_pairs() with the geom pairs MuJoCo would hand to collision_cb, add_flat_sensor() per
    FlatTactileSensor, finalize(), then step() every simulation step.
    """

    def __init__(self, n_envs, representation=REP_POLYGON, apply_contact_forces=True, device=0, max_faces=0,
                 max_candidates_per_slice=0, max_tactile_triangles=0, max_triangles_per_taxel=0, stream=None,
                 face_vertices=False):
        self.L = load_library()
        cfg = HcsConfig(device, n_envs, representation, int(apply_contact_forces), max_candidates_per_slice, max_faces,
                        max_tactile_triangles, max_triangles_per_taxel, stream, int(face_vertices))
        h = C.c_void_p()
        st = self.L.hcs_create(C.byref(cfg), C.byref(h))
        if st != HCS_OK:
            raise HcsError(st, self.L.hcs_last_error(None).decode())
        self.h = h
        self.n_envs = n_envs
        self.device = device
        self.sensors = []

    def close(self):
        if getattr(self, "h", None):
            self.L.hcs_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, st):
        if st < 0:
            raise HcsError(st, self.L.hcs_last_error(self.h).decode())
        return st

    # --- configuration ------------------------------------------------------------------------------
    def add_geom(self, mj_type, size, props, mesh_vert=None, mesh_face=None):
        size = _f64(np.resize(np.asarray(size, dtype=np.float64), 3))
        props = _f64(props)
        mv = np.ascontiguousarray(mesh_vert, dtype=np.float32) if mesh_vert is not None else None
        mf = np.ascontiguousarray(mesh_face, dtype=np.int32) if mesh_face is not None else None
        return self._check(self.L.hcs_add_geom(self.h, int(mj_type), _ptr(size, C.c_double), _ptr(mv, C.c_float),
                                               0 if mv is None else len(mv), _ptr(mf, C.c_int32),
                                               0 if mf is None else len(mf), _ptr(props, C.c_double)))

    def add_soft_mesh(self, verts, tets, pressure, props):
        verts, pressure, props = _f64(verts), _f64(pressure), _f64(props)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        return self._check(self.L.hcs_add_soft_mesh(self.h, _ptr(verts, C.c_double), len(verts), _ptr(tets, C.c_int32),
                                                    len(tets), _ptr(pressure, C.c_double), _ptr(props, C.c_double)))

    def add_rigid_mesh(self, verts, tris, props):
        verts, props = _f64(verts), _f64(props)
        tris = np.ascontiguousarray(tris, dtype=np.int32)
        return self._check(self.L.hcs_add_rigid_mesh(self.h, _ptr(verts, C.c_double), len(verts), _ptr(tris, C.c_int32),
                                                     len(tris), _ptr(props, C.c_double)))

    def update_geom(self, geom, size):
        size = _f64(np.resize(np.asarray(size, dtype=np.float64), 3))
        self._check(self.L.hcs_update_geom(self.h, int(geom), _ptr(size, C.c_double)))

    def set_env_sizes(self, geom, sizes):
        """Per-environment sizes [n_envs][3] of one geom (hcs_set_env_sizes); None: one size for all again."""
        if sizes is None:
            self._check(self.L.hcs_set_env_sizes(self.h, int(geom), None))
            return
        sizes = _f64(np.asarray(sizes, dtype=np.float64).reshape(self.n_envs, 3))
        self._check(self.L.hcs_set_env_sizes(self.h, int(geom), sizes.ctypes.data))

    def set_pairs(self, pairs):
        pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        g1, g2 = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
        self._check(self.L.hcs_set_pairs(self.h, _ptr(g1, C.c_int32), _ptr(g2, C.c_int32), len(pairs)))

    def add_flat_sensor(self, geom, resolution, sampling_resolution, window=WINDOW_NONE, sigma=-1.0):
        s = self._check(self.L.hcs_add_flat_sensor(self.h, int(geom), float(resolution), int(sampling_resolution),
                                                   int(window), float(sigma)))
        cx, cy = C.c_int(), C.c_int()
        self._check(self.L.hcs_sensor_dims(self.h, s, C.byref(cx), C.byref(cy)))
        self.sensors.append((cx.value, cy.value))
        return s

    def update_flat_sensor(self, sensor, sampling_resolution, window=WINDOW_NONE, sigma=-1.0):
        """FlatTactileSensor::dynamicParamCallback: new sampling_resolution / window / sigma for an existing sensor."""
        self._check(self.L.hcs_update_flat_sensor(self.h, int(sensor), int(sampling_resolution), int(window),
                                                  C.c_float(sigma)))

    def finalize(self):
        self._check(self.L.hcs_finalize(self.h))
        self.n_geoms = self.L.hcs_n_geoms(self.h)
        self.n_pairs = self.L.hcs_n_pairs(self.h)

    # --- stepping -------------------------------------------------------------------------------------
    def step(self, xpos, xmat, vel=None, with_sensors=False):
        """End-to-end entry point with HOST arrays [n_envs][n_geoms][3|9|6]; copies are part of the call."""
        n = self.n_envs * self.n_geoms
        xpos, xmat = _f64(xpos), _f64(xmat)
        vel = np.zeros(n * 6) if vel is None else _f64(vel)
        assert xpos.size == n * 3 and xmat.size == n * 9 and vel.size == n * 6, "pose arrays have the wrong size"
        self._check(self.L.hcs_step(self.h, xpos.ctypes.data, xmat.ctypes.data, vel.ctypes.data, int(with_sensors)))

    def step_raw(self, xpos_ptr, xmat_ptr, vel_ptr, with_sensors=False):
        """hcs_step with raw host addresses (e.g. pinned torch tensors)."""
        self._check(self.L.hcs_step(self.h, xpos_ptr, xmat_ptr, vel_ptr, int(with_sensors)))

    def step_async(self, xpos_ptr, xmat_ptr, vel_ptr, with_sensors=False, geom_wrench_ptr=None, sensor_image_ptrs=None,
                   pair_results_ptr=None):
        """hcs_step_async with raw HOST addresses (pinned memory keeps the copies asynchronous): queues the step and
        returns its ticket; results land in the caller's buffers (wait(ticket) tells when)."""
        out = HcsOutputs()
        out.geom_wrench = geom_wrench_ptr
        out.pair_results = pair_results_ptr
        keep = None
        if sensor_image_ptrs:
            keep = (C.c_void_p * len(sensor_image_ptrs))(*sensor_image_ptrs)
            out.sensor_images = C.cast(keep, C.POINTER(C.c_void_p))
        t = C.c_int64(-1)
        self._check(self.L.hcs_step_async(self.h, xpos_ptr, xmat_ptr, vel_ptr, int(with_sensors), C.byref(out), C.byref(t)))
        return t.value

    def wait(self, ticket):
        self._check(self.L.hcs_wait(self.h, C.c_int64(ticket)))

    def step_device(self, xpos_ptr, xmat_ptr, vel_ptr, with_sensors=False):
        """Asynchronous step on device pointers (ints, e.g. torch.Tensor.data_ptr())."""
        self._check(self.L.hcs_step_device(self.h, xpos_ptr, xmat_ptr, vel_ptr, int(with_sensors)))

    def sync(self):
        self._check(self.L.hcs_sync(self.h))

    def fetch(self, with_sensors=False):
        self._check(self.L.hcs_fetch_results(self.h, int(with_sensors)))

    # --- results ----------------------------------------------------------------------------------------
    def pair_results(self):
        out = np.zeros((self.n_envs, self.n_pairs), dtype=PAIR_RESULT_DTYPE)
        self._check(self.L.hcs_get_pair_results(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def geom_wrenches(self):
        out = np.zeros((self.n_envs, self.n_geoms, 6))
        self._check(self.L.hcs_get_geom_wrenches(self.h, _ptr(out, C.c_double)))
        return out

    def sensor_image(self, sensor):
        cx, cy = self.sensors[sensor]
        out = np.zeros((self.n_envs, cx * cy), dtype=np.float32)
        self._check(self.L.hcs_get_sensor_image(self.h, int(sensor), _ptr(out, C.c_float)))
        return out

    def add_curved_sensor(self, geom, taxel_pos, taxel_nrm, sample_pos, sample_nrm, include_margin):
        """CurvedSensor with caller-supplied surface samples (geom frame); returns the curved sensor index."""
        tp, sp, sn = _f64(taxel_pos).reshape(-1, 3), _f64(sample_pos).reshape(-1, 3), _f64(sample_nrm).reshape(-1, 3)
        tn = None if taxel_nrm is None else _f64(taxel_nrm).reshape(-1, 3)
        s = self._check(self.L.hcs_add_curved_sensor(self.h, int(geom), len(tp), tp.ctypes.data,
                                                     None if tn is None else tn.ctypes.data, len(sp), sp.ctypes.data,
                                                     sn.ctypes.data, float(include_margin)))
        if not hasattr(self, "curved"):
            self.curved = []
        self.curved.append(len(tp))
        return s

    def add_taxel_sensor(self, geom, taxel_pos, include_margin, sample_resolution, method="squared", visualize=False,
                         sample_method="default"):
        """TaxelSensor; method: closest | weighted | mean | squared; sample_method: default | area_importance."""
        tp = _f64(taxel_pos).reshape(-1, 3)
        code = {"closest": 0, "weighted": 1, "mean": 2, "squared": 3}[method]
        sm = {"default": 0, "area_importance": 1}[sample_method]
        s = self._check(self.L.hcs_add_taxel_sensor(self.h, int(geom), len(tp), tp.ctypes.data, float(include_margin),
                                                    float(sample_resolution), code, int(visualize), sm))
        if not hasattr(self, "taxel"):
            self.taxel = []
        self.taxel.append(len(tp))
        return s

    def taxel_values(self, sensor):
        out = np.zeros((self.n_envs, self.taxel[sensor]), dtype=np.float32)
        self._check(self.L.hcs_get_taxel_values(self.h, int(sensor), _ptr(out, C.c_float)))
        return out

    def curved_info(self, sensor):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.hcs_curved_sensor_info(self.h, int(sensor), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value  # taxels, rays (distinct assigned samples), (taxel, sample) assignments

    def curved_values(self, sensor):
        out = np.zeros((self.n_envs, self.curved[sensor]), dtype=np.float32)
        self._check(self.L.hcs_get_curved_values(self.h, int(sensor), _ptr(out, C.c_float)))
        return out

    def faces(self, cap=1 << 20):
        buf = np.zeros(cap, dtype=FACE_DTYPE)
        n = self._check(self.L.hcs_get_faces(self.h, buf.ctypes.data_as(C.c_void_p), cap))
        return buf[:min(n, cap)]

    def face_vertices(self, cap=1 << 20):
        """(n_faces, 8, 3) world vertices of the faces of faces(), same order; rows beyond a face's vertex count are 0"""
        n = self._check(self.L.hcs_get_faces(self.h, None, 0))
        n = min(n, cap)
        buf = np.zeros((max(n, 1), 8, 3))
        self._check(self.L.hcs_get_face_vertices(self.h, _ptr(buf, C.c_double), n))
        return buf[:n]

    def emitted(self, env, pair):
        n = self._check(self.L.hcs_get_emitted(self.h, int(env), int(pair), None, 0))
        buf = np.zeros((max(n, 1), 3), dtype=np.int32)
        self._check(self.L.hcs_get_emitted(self.h, int(env), int(pair), _ptr(buf, C.c_int32), n))
        return buf[:n]

    def tactile_triangles(self, env):
        n = self._check(self.L.hcs_get_tactile_triangles(self.h, int(env), None, 0))
        buf = np.zeros((max(n, 1), 12))
        self._check(self.L.hcs_get_tactile_triangles(self.h, int(env), _ptr(buf, C.c_double), n))
        return buf[:n]

    def geom_mesh(self, geom):
        info = (C.c_int * 3)()
        self._check(self.L.hcs_geom_info(self.h, int(geom), info))
        kind, nv, ne = info[0], info[1], info[2]
        if kind == 2:
            return dict(kind=kind)
        per = 4 if kind == 1 else 3
        verts, elems, grad = np.zeros((nv, 3)), np.zeros((ne, per), dtype=np.int32), np.zeros((ne, 3))
        pressure = np.zeros(nv) if kind == 1 else None
        e0 = np.zeros(ne) if kind == 1 else None
        self._check(self.L.hcs_get_mesh(self.h, int(geom), _ptr(verts, C.c_double), _ptr(elems, C.c_int32),
                                        _ptr(pressure, C.c_double), _ptr(grad, C.c_double), _ptr(e0, C.c_double)))
        out = dict(kind=kind, verts=verts, elems=elems)
        if kind == 1:
            out.update(pressure=pressure, grad=grad, e0=e0)
        else:
            out.update(normal=grad)
        return out

    def lbvh(self, geom):
        """The soft geom's GPU-resident LBVH as a structured array of 64-byte nodes."""
        dt = np.dtype([("llo", "<f4", 3), ("lhi", "<f4", 3), ("rlo", "<f4", 3), ("rhi", "<f4", 3), ("left", "<i4"),
                       ("right", "<i4"), ("pad", "<f4", 2)])
        n = self._check(self.L.hcs_get_lbvh(self.h, int(geom), None, 0))
        out = np.zeros(n, dtype=dt)
        self._check(self.L.hcs_get_lbvh(self.h, int(geom), out.ctypes.data_as(C.c_void_p), n))
        return out

    def counters(self):
        out = (C.c_int64 * 5)()
        self._check(self.L.hcs_get_counters(self.h, out))
        return dict(candidates=out[0], polygons=out[1], faces=out[2], tactile_triangles=out[3], kernels=out[4])

    def set_profiling(self, enable=True):
        self._check(self.L.hcs_set_profiling(self.h, int(enable)))

    def stage_ms(self):
        out = (C.c_float * 7)()
        self._check(self.L.hcs_get_stage_ms(self.h, out))
        names = ["setup", "broadphase", "narrowphase", "reduce", "tactile", "unused", "total"]
        return dict(zip(names, [float(x) for x in out]))

    def device_pair_results_ptr(self):
        return self.L.hcs_device_pair_results(self.h)

    def device_sensor_image_ptr(self, sensor):
        return self.L.hcs_device_sensor_image(self.h, int(sensor))


class MultiDeviceEngine:
    """hcs_multi: ONE context spanning several GPUs (include/hcs.h), driven from one process and one Python thread; the
    library shards the environments by index over the devices and runs the blocks on its own C++ threads.  Same
    configuration / step surface as HydroelasticEngine (scenes.configure works on it)."""

    def __init__(self, n_envs, devices, representation=REP_POLYGON, apply_contact_forces=True, **kw):
        self.L = load_library()
        cfg = HcsConfig(device=0, n_envs=int(n_envs), representation=int(representation),
                        apply_contact_forces=int(apply_contact_forces),
                        max_candidates_per_slice=int(kw.get("max_candidates_per_slice", 0)), max_faces=int(kw.get("max_faces", 0)),
                        max_tactile_triangles=int(kw.get("max_tactile_triangles", 0)),
                        max_triangles_per_taxel=int(kw.get("max_triangles_per_taxel", 0)), stream=None, face_vertices=0)
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        st = self.L.hcs_multi_create(C.byref(cfg), dev, len(devices), C.byref(h))
        if st < 0:
            raise HcsError(st, self.L.hcs_multi_last_error(None).decode())
        self.h, self.n_envs, self.devices, self.sensors = h, int(n_envs), list(devices), []
        self.n_geoms = self.n_pairs = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.hcs_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, st):
        if st < 0:
            raise HcsError(st, self.L.hcs_multi_last_error(self.h).decode())
        return st

    def blocks(self):
        out = []
        for k in range(self.L.hcs_multi_n_blocks(self.h)):
            a, b, c = C.c_int(), C.c_int(), C.c_void_p()
            self._check(self.L.hcs_multi_block(self.h, k, C.byref(a), C.byref(b), C.byref(c)))
            out.append((a.value, b.value))
        return out

    def add_geom(self, mj_type, size, props, mesh_vert=None, mesh_face=None):
        size = _f64(np.resize(np.asarray(size, dtype=np.float64), 3))
        props = _f64(props)
        mv = np.ascontiguousarray(mesh_vert, dtype=np.float32) if mesh_vert is not None else None
        mf = np.ascontiguousarray(mesh_face, dtype=np.int32) if mesh_face is not None else None
        g = self._check(self.L.hcs_multi_add_geom(self.h, int(mj_type), _ptr(size, C.c_double), _ptr(mv, C.c_float),
                                                  0 if mv is None else len(mv), _ptr(mf, C.c_int32),
                                                  0 if mf is None else len(mf), _ptr(props, C.c_double)))
        self.n_geoms = max(self.n_geoms, g + 1)
        return g

    def update_geom(self, geom, size):
        size = _f64(np.resize(np.asarray(size, dtype=np.float64), 3))
        self._check(self.L.hcs_multi_update_geom(self.h, int(geom), _ptr(size, C.c_double)))

    def set_pairs(self, pairs):
        pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        g1, g2 = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
        self._check(self.L.hcs_multi_set_pairs(self.h, _ptr(g1, C.c_int32), _ptr(g2, C.c_int32), len(pairs)))
        self.n_pairs = len(pairs)

    def add_flat_sensor(self, geom, resolution, sampling_resolution, window=WINDOW_NONE, sigma=-1.0):
        s = self._check(self.L.hcs_multi_add_flat_sensor(self.h, int(geom), float(resolution), int(sampling_resolution),
                                                         int(window), float(sigma)))
        a, b, c = C.c_int(), C.c_int(), C.c_void_p()
        self._check(self.L.hcs_multi_block(self.h, 0, C.byref(a), C.byref(b), C.byref(c)))
        cx, cy = C.c_int(), C.c_int()
        self.L.hcs_sensor_dims(c, s, C.byref(cx), C.byref(cy))
        self.sensors.append((cx.value, cy.value))
        return s

    def finalize(self):
        self._check(self.L.hcs_multi_finalize(self.h))

    def step(self, xpos, xmat, vel, with_sensors=False):
        xpos, xmat, vel = _f64(xpos), _f64(xmat), _f64(vel)
        n = self.n_envs * self.n_geoms
        assert xpos.size == n * 3 and xmat.size == n * 9 and vel.size == n * 6, "pose arrays have the wrong size"
        self._check(self.L.hcs_multi_step(self.h, xpos.ctypes.data, xmat.ctypes.data, vel.ctypes.data, int(with_sensors)))

    def step_async(self, xpos_ptr, xmat_ptr, vel_ptr, with_sensors=False, geom_wrench_ptr=None, sensor_image_ptrs=None,
                   pair_results_ptr=None):
        out = HcsOutputs()
        out.geom_wrench = geom_wrench_ptr
        out.pair_results = pair_results_ptr
        keep = None
        if sensor_image_ptrs:
            keep = (C.c_void_p * len(sensor_image_ptrs))(*sensor_image_ptrs)
            out.sensor_images = C.cast(keep, C.POINTER(C.c_void_p))
        t = C.c_int64(-1)
        self._check(self.L.hcs_multi_step_async(self.h, xpos_ptr, xmat_ptr, vel_ptr, int(with_sensors), C.byref(out), C.byref(t)))
        return t.value

    def wait(self, ticket):
        self._check(self.L.hcs_multi_wait(self.h, C.c_int64(ticket)))

    def geom_wrenches(self):
        out = np.zeros((self.n_envs, self.n_geoms, 6))
        self._check(self.L.hcs_multi_get_geom_wrenches(self.h, out.ctypes.data))
        return out

    def pair_results(self):
        out = np.zeros((self.n_envs, self.n_pairs), dtype=PAIR_RESULT_DTYPE)
        self._check(self.L.hcs_multi_get_pair_results(self.h, out.ctypes.data))
        return out

    def sensor_image(self, sensor):
        cx, cy = self.sensors[sensor]
        out = np.zeros((self.n_envs, cx * cy), dtype=np.float32)
        self._check(self.L.hcs_multi_get_sensor_image(self.h, int(sensor), out.ctypes.data))
        return out


def version():
    return load_library().hcs_version().decode()
