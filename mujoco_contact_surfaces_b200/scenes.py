"""Synthetic scenes of the shapes BASELINE.json names, engine-agnostic.

A scene is plain data: geoms in CONFIGURATION ORDER (the order of the `cs::<geom>` numerics in the
MuJoCo XML = drake_id order), the geom pairs MuJoCo's collision pass would hand to collision_cb, flat
sensors, and a seeded pose generator.  `configure()` feeds the same call sequence to any object with the
add_geom / set_pairs / add_flat_sensor surface (the CUDA engine, or the CPU oracle inside tests/bench).
MuJoCo itself is absent in this environment, so poses come from these generators instead of mj_step.
"""
import os

import numpy as np

GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)

_FIXTURES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class Geom:
    def __init__(self, name, mj_type, size, props, mesh_vert=None, mesh_face=None):
        self.name, self.mj_type = name, mj_type
        self.size = np.resize(np.asarray(size, dtype=np.float64), 3)
        self.props = np.asarray(props, dtype=np.float64)  # [E, dissipation, hint, mu_s, mu_d]
        self.mesh_vert, self.mesh_face = mesh_vert, mesh_face


class Scene:
    def __init__(self, name, geoms, pairs, triangle=False, sensors=(), apply_forces=True):
        self.name, self.geoms, self.pairs = name, geoms, [tuple(p) for p in pairs]
        self.triangle, self.sensors, self.apply_forces = triangle, list(sensors), apply_forces
        self.pose_fn = None
        self.env_sizes_fn = None  # optional: (n_envs, seed, env_offset) -> {geom index: sizes[n_envs][3]} (domain randomisation)
        # sizing hints for engine pools that cannot be derived from the geometry alone, per environment
        # (e.g. {"tactile_triangles_per_env": n}); engine_kwargs() scales them to a batch
        self.hints = {}

    def engine_kwargs(self, n_envs):
        kw = {}
        if "tactile_triangles_per_env" in self.hints:
            kw["max_tactile_triangles"] = int(self.hints["tactile_triangles_per_env"]) * n_envs
        return kw

    @property
    def n_geoms(self):
        return len(self.geoms)

    def env_sizes(self, n_envs, seed, env_offset=0):
        """Per-environment geom sizes for hcs_set_env_sizes, or {} when the scene has none."""
        return self.env_sizes_fn(n_envs, seed, env_offset) if self.env_sizes_fn else {}

    def poses(self, n_envs, seed, env_offset=0):
        """xpos[n,ng,3], xmat[n,ng,9], vel[n,ng,6] for envs env_offset .. env_offset+n-1 (per-env RNG streams,
        so a shard of envs sees exactly the poses the full batch would)."""
        ng = self.n_geoms
        xpos, xmat, vel = np.zeros((n_envs, ng, 3)), np.zeros((n_envs, ng, 9)), np.zeros((n_envs, ng, 6))
        for e in range(n_envs):
            rng = np.random.Generator(np.random.PCG64([seed, env_offset + e]))
            self.pose_fn(rng, env_offset + e, xpos[e], xmat[e], vel[e])
        return xpos, xmat, vel


def configure(target, scene):
    """Replay the scene's configuration calls on an engine-like object; returns its geom indices."""
    ids = []
    for g in scene.geoms:
        ids.append(target.add_geom(g.mj_type, g.size, g.props, g.mesh_vert, g.mesh_face))
    target.set_pairs([(ids[a], ids[b]) for a, b in scene.pairs])
    for s in scene.sensors:
        kw = dict(resolution=s["resolution"], sampling_resolution=s["sampling_resolution"],
                  window=s.get("window", 0), sigma=s.get("sigma", -1.0))
        if hasattr(target, "orc_needs_geom_size") or target.__class__.__name__ == "OracleScene":
            target.add_flat_sensor(ids[s["geom"]], scene.geoms[s["geom"]].size, **kw)
        else:
            target.add_flat_sensor(ids[s["geom"]], **kw)
    for s in getattr(scene, "taxel_sensors", []):
        target.add_taxel_sensor(ids[s["geom"]], s["taxel_pos"], s["include_margin"], s["sample_resolution"],
                                s.get("method", "squared"), s.get("visualize", False), s.get("sample_method", "default"))
    for s in getattr(scene, "curved_sensors", []):
        target.add_curved_sensor(ids[s["geom"]], s["taxel_pos"], s.get("taxel_nrm"), s["sample_pos"], s["sample_nrm"],
                                 s["include_margin"])
    return ids


# ---- helpers ------------------------------------------------------------------------------------------
def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rot_zyx(yaw, pitch, roll):
    cz, sz, cy, sy, cx, sx = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return Rz @ Ry @ Rx


def random_velocity(rng, vmax, wmax):
    def ball(r):
        d = rng.normal(size=3)
        return d / np.linalg.norm(d) * r * rng.uniform() ** (1 / 3)
    return np.concatenate([ball(wmax), ball(vmax)])  # (omega, v) like mj_objectVelocity


def load_mesh_fixture(name):
    """Meshes of the reference's example worlds, stored as de-duplicated float32 vertex / int32 face
    arrays (tests/golden/make_mesh_fixtures.py generated them from the reference assets)."""
    d = np.load(os.path.join(_FIXTURES, "meshes.npz"))
    return d[name + "_vert"].astype(np.float32), d[name + "_face"].astype(np.int32)


def box_surface_vertices(half, hint):
    n = [max(1, int(np.ceil(2 * h / hint))) for h in half]
    axes = [np.linspace(-h, h, k + 1) for h, k in zip(half, n)]
    g = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    return g


# ---- C1: soft sphere on rigid box (CS/assets/sphere_on_box_world.xml:22-41) ----------------------------
def sphere_on_box(identity_orientation=False):
    geoms = [Geom("box0", GEOM_BOX, [0.1, 0.1, 0.1], [0, 1.0, 0.1, 0.3, 0.3]),
             Geom("sphere0", GEOM_SPHERE, [0.08], [5e4, 5.0, 0.05, 0.3, 0.3])]
    # mj_collideGeoms orders a pair by geom type (sphere=2 < box=6): collision_cb(g1=sphere0, g2=box0)
    sc = Scene("c1_sphere_on_box", geoms, [(1, 0)], triangle=False)
    depths = [0.001, 0.005, 0.012, 0.03]

    def pose(rng, env, xpos, xmat, vel):
        d = depths[env % 4]
        dx, dy = rng.uniform(-0.05, 0.05, size=2)
        xpos[0] = [0, 0, 0.1]
        xmat[0] = np.eye(3).reshape(-1)
        xpos[1] = [dx, dy, 0.2 + 0.08 - d]
        R = np.eye(3) if identity_orientation else random_rotation(rng)
        xmat[1] = R.reshape(-1)
        vel[1] = random_velocity(rng, 0.1, 1.0)

    sc.pose_fn = pose
    return sc


# ---- C2: Myrmex foam pressed by a rigid object, 16x16 taxel image (SENS/assets/myrmex_*_world.xml) ------
def myrmex(presser="box", sampling_resolution=20, window=0, sigma=-1.0, resolution=0.025):
    if presser == "box":
        pg = Geom("box1", GEOM_BOX, [0.1, 0.1, 0.1], [0, 1.0, 0.05, 0.3, 0.3])
        pverts = box_surface_vertices([0.1, 0.1, 0.1], 0.05)
    elif presser == "plate":
        v, f = load_mesh_fixture("plate")
        pg, pverts = Geom("plate", GEOM_MESH, [0, 0, 0], [0, 1.0, 0, 0.3, 0.3], v, f), v.astype(np.float64)
    elif presser == "spot":
        v, f = load_mesh_fixture("spot")
        v = (v * np.float32(0.05)).astype(np.float32)
        pg, pverts = Geom("spot", GEOM_MESH, [0, 0, 0], [0, 1.0, 0, 0.3, 0.3], v, f), v.astype(np.float64)
    elif presser == "soft_tip":
        # C2b (BASELINE.json words config 2 as "pressed by soft mesh objects"): a SOFT convex-mesh presser,
        # the fingertip mesh of SENS/assets (ubi_tip_collision.stl, millimetres) four times life size so that it covers
        # several 25 mm taxels; centroid-fan tets (plugin.cpp:161-187, 767-787) => the soft-soft query (a7) feeds the sensor
        v, f = load_mesh_fixture("ubi_tip")
        v = (v * np.float32(0.004)).astype(np.float32)
        pg, pverts = Geom("soft_tip", GEOM_MESH, [0, 0, 0], [1e5, 2.0, 0, 0.6, 0.6], v, f), v.astype(np.float64)
    else:
        raise ValueError(presser)
    foam = Geom("myrmex_foam", GEOM_BOX, [0.2, 0.2, 0.02], [5e4, 5.0, 0, 0.3, 0.3])
    sensors = [dict(geom=1, resolution=resolution, sampling_resolution=sampling_resolution, window=window, sigma=sigma)]
    sc = Scene(("c2b_myrmex_" if presser.startswith("soft") else "c2_myrmex_") + presser, [pg, foam], [(0, 1)], triangle=True,
               sensors=sensors)
    foam_top = 0.033 + 0.02

    def pose(rng, env, xpos, xmat, vel):
        depth = rng.uniform(0.0005, 0.005)
        tilt = np.deg2rad(rng.uniform(-2, 2, size=2))
        yaw = rng.uniform(0, 2 * np.pi)
        R = rot_zyx(yaw, tilt[0], tilt[1])
        lowest = (pverts @ R.T)[:, 2].min()
        xy = rng.uniform(-0.05, 0.05, size=2)
        xpos[0] = [xy[0], xy[1], foam_top - depth - lowest]
        xmat[0] = R.reshape(-1)
        vel[0] = random_velocity(rng, 0.05, 0.5)
        xpos[1] = [0, 0, 0.033]
        xmat[1] = np.eye(3).reshape(-1)

    sc.pose_fn = pose
    return sc


def myrmex_multi(pressers=("box", "spot", "soft_tip"), sampling_resolution=8, resolution=0.025):
    """Several pressers on one Myrmex foam: one contact surface per presser, i.e. several BLAS under the sensor's TLAS
    (SENS/src/flat_tactile_sensor.cpp:285-302 builds one BVH per GeomCollision that touches the sensor geom)."""
    base = [myrmex(p, sampling_resolution, resolution=resolution) for p in pressers]
    geoms = [b.geoms[0] for b in base] + [base[0].geoms[1]]
    for i, g in enumerate(geoms[:-1]):
        g.name = "%s_%d" % (g.name, i)
    nf = len(geoms) - 1
    sensors = [dict(geom=nf, resolution=resolution, sampling_resolution=sampling_resolution, window=0, sigma=-1.0)]
    sc = Scene("c2_myrmex_multi", geoms, [(i, nf) for i in range(nf)], triangle=True, sensors=sensors)
    offsets = [np.array([0.11 * np.cos(a), 0.11 * np.sin(a)]) for a in np.linspace(0, 2 * np.pi, nf, endpoint=False)]

    def pose(rng, env, xpos, xmat, vel):
        xp, xm, ve = np.zeros((2, 3)), np.zeros((2, 9)), np.zeros((2, 6))
        for i, b in enumerate(base):
            b.pose_fn(rng, env, xp, xm, ve)
            xpos[i] = xp[0] * [0.3, 0.3, 1.0] + [offsets[i][0], offsets[i][1], 0.0]
            xmat[i], vel[i] = xm[0], ve[0]
        xpos[nf], xmat[nf] = xp[1], xm[1]

    sc.pose_fn = pose
    return sc


# ---- C3: two soft ellipsoids, equal-pressure-plane intersection -----------------------------------------
def soft_soft(hint=0.01, triangle=False):
    a, b = np.array([0.05, 0.04, 0.03]), np.array([0.04, 0.04, 0.06])
    geoms = [Geom("ell0", GEOM_ELLIPSOID, a, [5e4, 5.0, hint, 0.3, 0.3]),
             Geom("ell1", GEOM_ELLIPSOID, b, [1e5, 2.0, hint, 0.4, 0.2])]
    sc = Scene("c3_soft_soft", geoms, [(0, 1)], triangle=triangle)

    def support(axes, R, d):  # centre-to-surface distance of the rotated ellipsoid along direction d
        return 1.0 / np.linalg.norm((R.T @ d) / axes)

    def pose(rng, env, xpos, xmat, vel):
        R0, R1 = random_rotation(rng), random_rotation(rng)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        overlap = rng.uniform(0.001, 0.005)
        dist = support(a, R0, d) + support(b, R1, d) - overlap
        c0 = rng.uniform(-0.1, 0.1, size=3) + [0, 0, 0.3]
        xpos[0], xmat[0] = c0, R0.reshape(-1)
        xpos[1], xmat[1] = c0 + dist * d, R1.reshape(-1)
        vel[0], vel[1] = random_velocity(rng, 0.2, 2.0), random_velocity(rng, 0.2, 2.0)

    sc.pose_fn = pose
    return sc


# ---- C4: mixed soft objects dropped on a rigid plane ------------------------------------------------------
def objects_on_plane(triangle=False):
    tip_v, tip_f = load_mesh_fixture("ubi_tip")
    tip_v = (tip_v * np.float32(0.001)).astype(np.float32)  # the STL is in millimetres
    geoms = [Geom("ground", GEOM_PLANE, [0, 0, 1], [0, 1.0, 0, 0.5, 0.5]),
             Geom("sphere", GEOM_SPHERE, [0.05], [5e4, 5.0, 0.025, 0.3, 0.3]),
             Geom("ellipsoid", GEOM_ELLIPSOID, [0.06, 0.04, 0.03], [8e4, 3.0, 0.015, 0.3, 0.3]),
             Geom("box", GEOM_BOX, [0.05, 0.04, 0.03], [6e4, 4.0, 0, 0.3, 0.3]),
             Geom("tip", GEOM_MESH, [0, 0, 0], [1e5, 2.0, 0, 0.6, 0.6], tip_v, tip_f)]
    # plane (type 0) comes first in MuJoCo's pair order
    sc = Scene("c4_objects_on_plane", geoms, [(0, 1), (0, 2), (0, 3), (0, 4)], triangle=triangle)
    extents = [None, np.full(3, 0.05), np.array([0.06, 0.04, 0.03]), None, None]
    box_corners = np.array([[sx * 0.05, sy * 0.04, sz * 0.03] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    tip = tip_v.astype(np.float64)

    def env_sizes(n_envs, seed, env_offset=0):
        """Domain-randomised sizes (sizes drawn per environment): every object scaled by its own
        factor in [0.92, 1.08] (inside one refinement level of the sphere and the ellipsoid; the box keeps its proportions,
        hence its medial-axis topology)."""
        out = {1: np.zeros((n_envs, 3)), 2: np.zeros((n_envs, 3)), 3: np.zeros((n_envs, 3))}
        for e in range(n_envs):
            rng = np.random.Generator(np.random.PCG64([seed, 7777, env_offset + e]))
            f = rng.uniform(0.92, 1.08, size=3)
            out[1][e] = [0.05 * f[0], 0, 0]
            out[2][e] = np.array([0.06, 0.04, 0.03]) * f[1]
            out[3][e] = np.array([0.05, 0.04, 0.03]) * f[2]
        return out

    sc.env_sizes_fn = env_sizes

    def pose(rng, env, xpos, xmat, vel):
        xmat[0] = np.eye(3).reshape(-1)
        for g in range(1, 5):
            R = random_rotation(rng)
            if g in (1, 2):
                low, size = -np.linalg.norm(extents[g] * R[2]), 2 * extents[g].min()
            elif g == 3:
                low, size = (box_corners @ R.T)[:, 2].min(), 0.06
            else:
                z = (tip @ R.T)[:, 2]
                low, size = z.min(), z.max() - z.min()
            depth = rng.uniform(0, 0.2 * size)
            xpos[g] = [rng.uniform(-1, 1), rng.uniform(-1, 1), -low - depth]
            xmat[g] = R.reshape(-1)
            vel[g] = random_velocity(rng, 0.5, 5.0)

    sc.pose_fn = pose
    return sc


# ---- C5: multi-finger grasp, high-resolution fingertip pads with one flat tactile array each --------------------
def grasp(obj="box", pad_hint=0.00025, n_pads=5, sampling_resolution=8, pad_axes=(0.010, 0.010, 0.012)):
    """`n_pads` soft ellipsoid pads (single-interior-vertex mesh; pad_hint 0.00025 -> refinement level 7 =
    131072 tets, 65539 vertices) pressed 0.2-2 mm into a rigid object (cube of half size 0.03 with 3072
    triangles, or the spot mesh), one pad per face.  Each pad carries a FlatTactileSensor: the sensor frame is the
    pad's own frame (flat_tactile_sensor.cpp:290-340), so the pad's +z axis points at the object and the rays run
    from 1.5 zs beyond the pad centre back through the contact patch.  Equal x/y semi-axes give the 16 x 16 array
    BASELINE.json names (SURVEY.md 8d lists (0.010, 0.008, 0.012), which would be 16 x 12 taxels)."""
    pad_axes = np.asarray(pad_axes, dtype=np.float64)
    half = 0.03
    if obj == "box":
        og = Geom("object", GEOM_BOX, [half] * 3, [0, 1.0, 2 * half / 16, 0.8, 0.8])
        pad_first = True  # mj_collideGeoms orders by geom type: ellipsoid (4) < box (6)
    elif obj == "spot":
        v, f = load_mesh_fixture("spot")
        v = (v * np.float32(0.05)).astype(np.float32)
        og = Geom("object", GEOM_MESH, [0, 0, 0], [0, 1.0, 0, 0.8, 0.8], v, f)
        pad_first = True  # ellipsoid (4) < mesh (7)
    else:
        raise ValueError(obj)
    geoms = [og] + [Geom("pad%d" % i, GEOM_ELLIPSOID, pad_axes, [5e4, 5.0, pad_hint, 0.8, 0.8]) for i in range(n_pads)]
    pairs = [(1 + i, 0) if pad_first else (0, 1 + i) for i in range(n_pads)]
    sensors = [dict(geom=1 + i, resolution=2 * pad_axes[0] / 16, sampling_resolution=sampling_resolution)
               for i in range(n_pads)]
    sc = Scene("c5_grasp_" + obj, geoms, pairs, triangle=True, sensors=sensors)
    normals = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    spot_v = og.mesh_vert.astype(np.float64) if obj == "spot" else None
    if spot_v is not None:
        tri0, tri1, tri2 = (spot_v[og.mesh_face[:, k]] for k in range(3))

    def mesh_height(o, n):
        """Largest s with o + s n on the mesh surface (the point a finger moving along -n touches first)."""
        e1, e2 = tri1 - tri0, tri2 - tri0
        pv = np.cross(n, e2)
        det = np.einsum("ij,ij->i", e1, pv)
        ok = np.abs(det) > 1e-14
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = o - tri0
        u = np.einsum("ij,ij->i", tv, pv) * inv
        qv = np.cross(tv, e1)
        v = (qv @ n) * inv
        s = np.einsum("ij,ij->i", qv, e2) * inv
        hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1)
        return s[hit].max() if hit.any() else None

    def pose(rng, env, xpos, xmat, vel):
        R_W, p_W = random_rotation(rng), rng.uniform(-0.2, 0.2, size=3) + [0, 0, 0.5]
        v_obj = random_velocity(rng, 0.05, 0.5)
        xpos[0], xmat[0], vel[0] = p_W, R_W.reshape(-1), v_obj
        for i in range(n_pads):
            n = normals[i % 6]
            t1 = np.roll(n, 1)  # two tangents of the face
            t2 = np.cross(n, t1)
            uv = rng.uniform(-0.4, 0.4, size=2) * half
            if spot_v is None:
                surf = half
            else:  # first surface point met along -n at the chosen tangential offset (the mesh spans ~ +-0.03)
                surf = mesh_height(uv[0] * t1 + uv[1] * t2, n)
                if surf is None:
                    uv = np.zeros(2)
                    surf = mesh_height(np.zeros(3), n)
            depth = rng.uniform(0.0002, 0.002)
            yaw = rng.uniform(0, 2 * np.pi)
            tilt = np.deg2rad(rng.uniform(-3, 3, size=2))
            # pad frame in the object frame: local +z = -n (towards the object), then yaw / tilt about it
            Rz = np.column_stack([t1, -t2, -n])
            R_OP = Rz @ rot_zyx(yaw, tilt[0], tilt[1])
            c_O = uv[0] * t1 + uv[1] * t2 + (surf + pad_axes[2] - depth) * n
            xpos[1 + i] = p_W + R_W @ c_O
            xmat[1 + i] = (R_W @ R_OP).reshape(-1)
            v_rel = random_velocity(rng, 0.01, 0.1)
            vel[1 + i] = v_obj + v_rel

    # fan triangles per pad and env: tets under the largest contact patch (2 mm deep) x 1.5 (polygons split
    # along the object's triangle edges) x 5 fan triangles per polygon
    level = max(0, int(np.ceil(np.log2(np.pi * pad_axes.max() / (2 * pad_hint)))))
    tet_area = 4 * np.pi * pad_axes.mean() ** 2 / (8 * 4 ** level)
    # (the spot mesh has ~2 mm^2 triangles and concave regions: patches are larger and cut into more polygons)
    per_patch = (30.0 if obj == "spot" else 7.5) * 2 * np.pi * pad_axes.max() * 0.002 / tet_area
    sc.hints["tactile_triangles_per_env"] = int(n_pads * max(12000 if obj == "spot" else 2000, per_patch))
    sc.pose_fn = pose
    return sc


# ---- curved fingertip sensor (SENS/assets/fingertip_mocap.xml + SENS/config/curved_fingertip.yaml) ----------------
# taxel positions / normals of the reference's fingertip configuration (geom frame of the ubi_tip mesh, metres)
_TIP_TAXELS = np.array([
    [0.00415178164, -0.00615073064, 0.01464621212], [0.008619488, 0.00046286059, 0.01316570371],
    [0.00830223182, 0.00074051459, 0.01945107626], [0.00354169091, -0.00500616896, 0.0238568657],
    [0.00266992694, 0.00055709812, 0.03028623604], [0.00713842137, 0.00097836545, 0.02588591432],
    [-0.00713842136, 0.00097836546, 0.02588591435], [-0.00266992693, 0.00055709811, 0.030286236],
    [-0.00354169087, -0.00500616897, 0.02385686587], [-0.00830223175, 0.00074051458, 0.01945107617],
    [-0.00861948819, 0.00046286058, 0.01316570354], [-0.0041517816, -0.0061507306, 0.01464621205]])
_TIP_NORMALS = np.array([
    [0.37386554, -0.92331819, 0.08779566], [0.9989708, -0.03265917, 0.03147561], [0.99543803, -0.03717611, 0.08786959],
    [0.36603357, -0.89345384, 0.26030685], [0.35214134, -0.13784856, 0.92573984], [0.9230213, -0.06311386, 0.37953699],
    [-0.9230213, -0.06311386, 0.37953699], [-0.35214134, -0.13784856, 0.92573984], [-0.36603357, -0.89345384, 0.26030685],
    [-0.99543803, -0.03717611, 0.08786959], [-0.9989708, -0.03265917, 0.03147561], [-0.37386554, -0.92331819, 0.08779566]])


def surface_samples(verts, faces, n, seed=42):
    """Area-weighted random points on a triangle mesh with the face normals of the winding (stand-in for the
    reference's vcglib Poisson-disk sampling, curved_sensor.cpp:271-283, which is not reproducible here)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tri = verts[faces]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    pick = rng.choice(len(faces), size=n, p=area / area.sum())
    u, v = rng.uniform(size=n), rng.uniform(size=n)
    a = 1 - np.sqrt(u)
    b = (1 - a) * v
    pts = a[:, None] * tri[pick, 0] + (1 - a - b)[:, None] * tri[pick, 1] + b[:, None] * tri[pick, 2]
    return pts, nrm[pick] / (2 * area[pick])[:, None]


def myrmex_taxels(presser="box", method="weighted", visualize=False, sample_method="default", sample_resolution=0.01):
    """SENS/config/flat_taxel_sensor.yaml: the Myrmex foam read by a TaxelSensor with a 16 x 16 taxel lattice on the
    foam's top face (include_margin = half the taxel pitch, sample_resolution 0.01)."""
    sc = myrmex(presser, sampling_resolution=4)
    sc.name = "taxel_myrmex_" + presser
    g = np.linspace(-0.19, 0.19, 16)
    taxels = np.array([[x, y, 0.02] for x in g for y in g])
    sc.taxel_sensors = [dict(geom=1, taxel_pos=taxels, include_margin=0.01266666666666666,
                             sample_resolution=sample_resolution, method=method, visualize=visualize,
                             sample_method=sample_method)]
    return sc


def fingertip(n_samples=3000, include_margin=0.006, with_normals=True):
    """Soft ubi_tip fingertip (convex-mesh geom, centroid-fan tets) pressed onto a rigid box; CurvedSensor with the
    reference's 12 taxels on the fingertip geom."""
    tip_v, tip_f = load_mesh_fixture("ubi_tip")
    tip_v = (tip_v * np.float32(0.001)).astype(np.float32)  # fingertip_mocap.xml:42
    geoms = [Geom("box_geom", GEOM_BOX, [0.025] * 3, [0, 1.0, 0.01, 0.0, 0.0]),
             Geom("fingertip_geom", GEOM_MESH, [0, 0, 0], [5e4, 5.0, 0.0, 0.0, 0.0], tip_v, tip_f)]
    sc = Scene("fingertip_curved", geoms, [(0, 1)], triangle=True)
    tv = tip_v.astype(np.float64)
    pts, nrm = surface_samples(tv, tip_f, n_samples)
    sc.curved_sensors = [dict(geom=1, taxel_pos=_TIP_TAXELS, taxel_nrm=_TIP_NORMALS if with_normals else None,
                              sample_pos=pts, sample_nrm=nrm, include_margin=include_margin)]
    sc.hints["tactile_triangles_per_env"] = 20000

    def pose(rng, env, xpos, xmat, vel):
        xpos[0], xmat[0] = [0, 0, 0.025], np.eye(3).reshape(-1)
        n = _TIP_NORMALS[rng.integers(len(_TIP_NORMALS))]
        n = n / np.linalg.norm(n)
        # rotation taking the chosen taxel normal to -z, then yaw about z and a small tilt
        t = np.cross(n, [0.3, 0.5, 0.8])
        t /= np.linalg.norm(t)
        B = np.column_stack([t, np.cross(-n, t), -n])  # local frame whose third axis is -n
        R = rot_zyx(rng.uniform(0, 2 * np.pi), *np.deg2rad(rng.uniform(-8, 8, size=2))) @ B.T
        depth = rng.uniform(0.0003, 0.002)
        low = (tv @ R.T)[:, 2].min()
        xy = rng.uniform(-0.01, 0.01, size=2)
        xpos[1], xmat[1] = [xy[0], xy[1], 0.05 - depth - low], R.reshape(-1)
        vel[1] = random_velocity(rng, 0.02, 0.2)

    sc.pose_fn = pose
    return sc


SCENES = {
    "c1_sphere_on_box": sphere_on_box,
    "c2_myrmex_box": lambda: myrmex("box"),
    "c2_myrmex_plate": lambda: myrmex("plate"),
    "c2_myrmex_spot": lambda: myrmex("spot"),
    "c2b_myrmex_soft_tip": lambda: myrmex("soft_tip"),
    "c3_soft_soft": soft_soft,
    "c4_objects_on_plane": objects_on_plane,
    "c5_grasp_box": lambda: grasp("box"),
    "c5_grasp_spot": lambda: grasp("spot"),
}


# ---- every mesh family of plugin.cpp:633-808 in one scene (parity coverage, not a benchmark) ---------------
def mixed_shapes(triangle=False):
    geoms = [Geom("rigid_box", GEOM_BOX, [0.3, 0.3, 0.05], [0, 1.0, 0.1, 0.4, 0.4]),            # 0 rigid grid box
             Geom("soft_cyl_long", GEOM_CYLINDER, [0.03, 0.06, 0], [7e4, 3.0, 0.02, 0.3, 0.3]),   # 1 MA cylinder, segment
             Geom("soft_cyl_flat", GEOM_CYLINDER, [0.06, 0.02, 0], [7e4, 3.0, 0.03, 0.3, 0.3]),   # 2 MA cylinder, disc
             Geom("soft_box_grid", GEOM_BOX, [0.04, 0.05, 0.03], [4e4, 2.0, 0.03, 0.2, 0.2]),      # 3 soft grid box
             Geom("rigid_sphere", GEOM_SPHERE, [0.05], [0, 1.0, 0.02, 0.5, 0.5]),                  # 4 rigid sphere
             Geom("rigid_ellipsoid", GEOM_ELLIPSOID, [0.05, 0.03, 0.04], [0, 1.0, 0.02, 0.5, 0.5]),  # 5
             Geom("rigid_cyl", GEOM_CYLINDER, [0.04, 0.05, 0], [0, 1.0, 0.02, 0.5, 0.5]),          # 6 rigid cylinder
             Geom("soft_box_ma", GEOM_BOX, [0.05, 0.05, 0.05], [6e4, 4.0, 0, 0.3, 0.3])]           # 7 soft MA cube
    # soft things resting on the rigid box, rigid things pressed into soft things, one soft-soft pair
    pairs = [(1, 0), (2, 0), (0, 3), (4, 3), (5, 7), (6, 7), (1, 2)]
    sc = Scene("mixed_shapes", geoms, pairs, triangle=triangle)
    top = 0.05

    def pose(rng, env, xpos, xmat, vel):
        xmat[:] = np.eye(3).reshape(-1)
        xpos[0] = [0, 0, 0]
        # cylinders lying / standing on the box with small random tilt
        xpos[1] = [-0.15, -0.15, top + 0.06 - rng.uniform(0.002, 0.01)]
        xmat[1] = rot_zyx(rng.uniform(0, 6.28), *np.deg2rad(rng.uniform(-5, 5, size=2))).reshape(-1)
        xpos[2] = [-0.13, -0.13 + rng.uniform(-0.01, 0.01), top + 0.06 + 0.06 + 0.02 - rng.uniform(0.012, 0.02)]
        xmat[2] = rot_zyx(rng.uniform(0, 6.28), *np.deg2rad(rng.uniform(-3, 3, size=2))).reshape(-1)
        xpos[3] = [0.15, 0.15, top + 0.03 - rng.uniform(0.002, 0.008)]
        xmat[3] = rot_zyx(rng.uniform(0, 6.28), *np.deg2rad(rng.uniform(-4, 4, size=2))).reshape(-1)
        xpos[4] = [0.15 + rng.uniform(-0.01, 0.01), 0.15, xpos[3][2] + 0.03 + 0.05 - rng.uniform(0.003, 0.01)]
        xmat[4] = random_rotation(rng).reshape(-1)
        xpos[7] = [0.6, 0, 0.3]
        xmat[7] = random_rotation(rng).reshape(-1)
        d = random_rotation(rng)[:, 0]
        xpos[5] = xpos[7] + d * 0.075
        xmat[5] = random_rotation(rng).reshape(-1)
        xpos[6] = xpos[7] - d * 0.08
        xmat[6] = random_rotation(rng).reshape(-1)
        for g in (1, 2, 3, 4, 5, 6):
            vel[g] = random_velocity(rng, 0.2, 2.0)

    sc.pose_fn = pose
    return sc
