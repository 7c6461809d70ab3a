#!/usr/bin/env python
"""The reference's own flat-sensor benchmark grid on the CUDA engine.

mujoco_contact_surface_sensors/src/benchmark/benchmark_flat.cpp runs the Myrmex worlds {Box, Plate, Spot} x taxel
resolution {0.025, 0.0025} (16 x 16 / 160 x 160 taxels) x sampling_resolution {4, 8, 16, 32} x use_parallel {off, on}
with ONE mjData, 100 timed steps, and writes per-step times of the contact surface and of the sensor update to
/tmp/bench_0{1..4}.csv (columns :52: ..., surface_*, bblas, btlas, tr, total, nrays, ntri).  This script runs the same
grid through the C ABI (`hcs_step` with host buffers, one taxel image per step):

  * impl "gpu":          one environment (the reference's case) and a batch, wall clock of hcs_step + per-stage CUDA events
  * impl "cpu_serial" /  the CPU restatement of the reference path, use_parallel off / on (OpenMP over the taxels,
    "cpu_parallel":      flat_tactile_sensor.cpp:316), on a bounded number of updates; timed by `bench.py --cpu-flat`

  python benchmark_flat.py [--csv profiles/r01_benchmark_flat.csv] [--batch 64] [--steps 20] [--no-cpu]
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", default=os.path.join(ROOT, "profiles", "r01_benchmark_flat.csv"))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=2.0, help="CPU time budget per (config, mode)")
    args = ap.parse_args()

    from mujoco_contact_surfaces_b200 import REP_TRIANGLE, HydroelasticEngine
    from mujoco_contact_surfaces_b200 import scenes as S

    rows = []
    for resolution in (0.025, 0.0025):
        for sampling in (4, 8, 16, 32):
            for mesh in ("box", "plate", "spot"):
                scene = S.myrmex(mesh, sampling_resolution=sampling, resolution=resolution)
                cx = int(np.floor(2 * 0.2 / resolution + 0.1))
                nrays = cx * cx * sampling * sampling
                for n_envs in (1, args.batch):
                    # the taxel bins scale with the image: (triangle, taxel) overlaps per taxel stay small, but the
                    # fine grid has 100 x the taxels
                    eng = HydroelasticEngine(n_envs, representation=REP_TRIANGLE, **scene.engine_kwargs(n_envs))
                    S.configure(eng, scene)
                    eng.finalize()
                    sets = [scene.poses(n_envs, seed=7 + i) for i in range(4)]
                    for i in range(3):
                        eng.step(*sets[i % 4], with_sensors=True)
                    t0 = time.perf_counter()
                    for i in range(args.steps):
                        eng.step(*sets[i % 4], with_sensors=True)
                    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
                    eng.set_profiling(True)
                    stage = {}
                    for i in range(args.steps):
                        eng.step(*sets[i % 4], with_sensors=True)
                        for k, v in eng.stage_ms().items():
                            stage[k] = stage.get(k, 0.0) + v / args.steps
                    eng.set_profiling(False)
                    c = eng.counters()
                    img = eng.sensor_image(0)
                    assert img.shape == (n_envs, cx * cx) and img.max() > 0
                    rows.append(dict(impl="gpu", n_envs=n_envs, resolution=resolution, sampling_resolution=sampling, mesh=mesh,
                                     surface_ms=stage["broadphase"] + stage["narrowphase"] + stage["reduce"],
                                     sensor_ms=stage["tactile"], total_ms=wall_ms, ms_per_env=wall_ms / n_envs,
                                     nrays=nrays, ntri=c["tactile_triangles"] // n_envs))
                    print(rows[-1], flush=True)
                    eng.close()
                if args.no_cpu:
                    continue
                # the CPU legs run inside bench.py (the one place outside tests/ that may execute oracle/)
                out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--cpu-flat", mesh, str(resolution),
                                      str(sampling), str(args.cpu_seconds)], capture_output=True, text=True, check=True)
                for line in out.stdout.strip().splitlines():
                    r = json.loads(line)
                    r["nrays"] = nrays
                    rows.append(r)
                    print(r, flush=True)
    with open(args.csv, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    print("wrote", args.csv)


if __name__ == "__main__":
    main()
