"""CPU tests of the host logic: the C-ABI library loads and exports every declared symbol (no compute
without a GPU), scenes are reproducible, env sharding over world_size 2 (gloo) reproduces the
single-process batch."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest

from mujoco_contact_surfaces_b200 import scenes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_declared_in_the_header(hcs_lib):
    from mujoco_contact_surfaces_b200 import engine
    header = open(os.path.join(ROOT, "include", "hcs.h")).read()
    declared = set(re.findall(r"\b(hcs_[a-z_]+)\s*\(", header))
    declared -= {"hcs_ctx"}
    assert declared == set(engine.ABI_SYMBOLS), declared ^ set(engine.ABI_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(hcs_lib, name), "libhcs_b200.so does not export %s" % name
    assert b"sm_100a" in hcs_lib.hcs_version()


def test_struct_layouts_match_the_header(hcs_lib):
    from mujoco_contact_surfaces_b200 import engine
    assert engine.PAIR_RESULT_DTYPE.itemsize == 112  # 10 doubles + 8 int32
    assert engine.FACE_DTYPE.itemsize == 120  # 12 doubles + 6 int32
    assert ctypes.sizeof(engine.HcsConfig) == 8 * 4 + 8 + 8  # 8 ints, stream pointer, face_vertices + padding


def test_product_fails_loudly_without_a_gpu(hcs_lib):
    """No CPU fallback: on a box without CUDA devices hcs_create reports HCS_E_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mujoco_contact_surfaces_b200 import HcsError, HydroelasticEngine
    with pytest.raises(HcsError) as ei:
        HydroelasticEngine(4)
    assert ei.value.status == -3 and "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "mujoco_contact_surfaces_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_scene_poses_are_reproducible_and_shardable():
    sc = scenes.sphere_on_box()
    a = sc.poses(16, seed=7)
    b = sc.poses(16, seed=7)
    c = sc.poses(8, seed=7, env_offset=8)
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y) and np.array_equal(x[8:], z)
    R = a[1][:, 1].reshape(-1, 3, 3)
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-14)


def test_all_scenes_configure_on_the_oracle():
    from parity_utils import make_oracle
    for name, fn in scenes.SCENES.items():
        sc = fn()
        o = make_oracle(sc)
        xp, xm, ve = sc.poses(2, seed=3)
        o.step(xp[0], xm[0], ve[0])
        assert sum(o.pair_result(p)["n_polygons"] for p in range(len(sc.pairs))) > 0, name


def test_shard_ranges_partition_the_batch():
    for n, w in [(4096, 8), (10, 3), (5, 8), (1, 1)]:
        cover = []
        for r in range(w):
            s, c = sharding.shard_range(n, r, w)
            cover += list(range(s, s + c))
        assert cover == list(range(n))


class _OracleEngine:
    """Engine-shaped stand-in used ONLY by the CPU sharding test (the CUDA engine needs a GPU)."""

    def __init__(self, scene, n_envs):
        from mujoco_contact_surfaces_b200.engine import PAIR_RESULT_DTYPE
        from parity_utils import make_oracle
        self.scene, self.n_envs, self.o, self.dt = scene, n_envs, make_oracle(scene), PAIR_RESULT_DTYPE

    def step(self, xpos, xmat, vel, with_sensors=False):
        self.res = np.zeros((self.n_envs, len(self.scene.pairs)), dtype=self.dt)
        for e in range(self.n_envs):
            self.o.step(xpos[e], xmat[e], vel[e])
            for p in range(len(self.scene.pairs)):
                r = self.o.pair_result(p)
                self.res[e, p]["F"], self.res[e, p]["tau"] = r["F"], r["tau"]
                self.res[e, p]["n_polygons"] = r["n_polygons"]

    def pair_results(self):
        return self.res


def _worker(rank, world, port, n_total, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sc = scenes.sphere_on_box()
    batch = sharding.ShardedBatch(sc, n_total, lambda n: _OracleEngine(sc, n), rank, world)
    xp, xm, ve = batch.local_poses(seed=11)
    batch.step(xp, xm, ve)
    allres = batch.gather_pair_results()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), allres)
    dist.destroy_process_group()


def test_env_sharding_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    n_total = 7  # uneven split: 4 + 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    sc = scenes.sphere_on_box()
    single = _OracleEngine(sc, n_total)
    single.step(*sc.poses(n_total, seed=11))
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == single.res.shape
        assert got.tobytes() == single.res.tobytes()  # bit-identical to the unsharded batch


def _run_bench(*argv, env=None):
    import json
    import subprocess
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(argv), capture_output=True, text=True,
                         timeout=600, env=e)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(line) for line in out.stdout.strip().splitlines()]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours): one JSON line with the contract's keys,
    and every host thread in use even when the launcher exports OMP_NUM_THREADS=1 like torchrun does."""
    (line,) = _run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--envs", "64", env={"OMP_NUM_THREADS": "1"})
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["value"] > 0
    assert line["config"]["workload"] == "c1_sphere_on_box" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reference_arm_other_ranks_stay_silent():
    assert _run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_benchmark_flat_cpu_leg():
    """CPU leg of benchmark_flat.py (the grid of the reference's benchmark_flat.cpp): serial and use_parallel rows."""
    rows = _run_bench("--cpu-flat", "box", "0.025", "4", "0.05")
    assert [r["impl"] for r in rows] == ["cpu_serial", "cpu_parallel"]
    for r in rows:
        assert r["mesh"] == "box" and r["sampling_resolution"] == 4 and r["ntri"] > 0 and r["total_ms"] > r["sensor_ms"] > 0


def test_abi_header_is_plain_c():
    """include/hcs.h is the drop-in boundary: it must compile as C99 (and as C++) on its own."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "hcs.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", hdr])


def test_plain_c_example_links_against_the_abi(hcs_lib, tmp_path):
    """examples/sphere_on_box.c uses the library from C99 through include/hcs.h alone; without a CUDA device it must
    fail loudly in hcs_create (no CPU fallback behind the boundary either)."""
    import subprocess
    import torch
    exe = str(tmp_path / "sphere_on_box")
    libdir = os.path.join(ROOT, "mujoco_contact_surfaces_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "sphere_on_box.c"), "-L", libdir, "-lhcs_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is covered by the GPU tests")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 1 and "no usable CUDA device" in out.stderr
