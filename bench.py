#!/usr/bin/env python
"""Headline benchmark: batched hydroelastic contact-surface steps on B200 (BASELINE.json metric:
contact-surface pair-evals/s & env-steps/s vs the CPU path).

  python bench.py --gpus N --steps K --warmup W          # our CUDA engine; under torchrun for N > 1
  python bench.py --impl reference ...                   # the reference's CPU execution model, timed on
                                                         # the host cores through the oracle restatement
                                                         # (the reference itself needs Drake/MuJoCo/ROS,
                                                         # which cannot be built here: SURVEY.md §8c)

One "step" = one contact pass (poses in -> broadphase -> narrowphase -> per-pair wrench [-> taxel images])
over a batch of n_envs independent environments per GPU.  Envs shard by index across GPUs with no
collective on the step path (weak scaling: n_envs per GPU is fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per pair-eval, fp64 geometry mode (DESIGN.md "Roofline")
BYTES_PER_PAIR = {"soft_rigid": 232, "soft_plane": 132, "soft_soft": 264}
BYTES_PER_POLYGON = 80


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1_sphere_on_box")
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--pose-sets", type=int, default=8)
    ap.add_argument("--cpu-sample-envs", type=int, default=0, help="0 = automatic (about 10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-events", action="store_true",
                    help="diagnostic: leave the per-stage CUDA events out of the timed steps (no roofline then)")
    ap.add_argument("--sensors", type=int, default=-1, help="-1: on when the workload has sensors")
    return ap.parse_args()


def workload_kinds(scene):
    kinds = []
    for a, b in scene.pairs:
        ga, gb = scene.geoms[a], scene.geoms[b]
        sa, sb = ga.props[0] > 0, gb.props[0] > 0
        if sa and sb:
            kinds.append("soft_soft")
        elif not (sa or sb):
            kinds.append(None)
        else:
            other = gb if sa else ga
            kinds.append("soft_plane" if other.mj_type == 0 else "soft_rigid")
    return kinds


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(scene, seed, sample_envs, with_sensors, threads_all=True):
    """Time the oracle (restated reference path, BVH broadphase) on the host cores: 1 thread (the
    reference's execution model: one physics thread, OpenMP only inside the taxel loop) and all threads
    (generous batched-CPU variant, OpenMP over envs)."""
    from mujoco_contact_surfaces_b200 import scenes as S
    from oracle import oracle
    orc = oracle.OracleScene(scene.triangle, scene.apply_forces)
    S.configure(orc, scene)
    n_threads = max(oracle.num_threads(), len(os.sched_getaffinity(0)))  # OMP_NUM_THREADS=1 under torchrun
    # calibrate
    xp, xm, ve = scene.poses(8, seed)
    t8, _, _ = orc.bench(xp, xm, ve, use_bvh=True, with_sensors=with_sensors, threads=1)
    per_env = max(t8 / 8, 1e-7)
    if sample_envs <= 0:
        sample_envs = int(min(4096, max(32, 2.0 / per_env)))
    xp, xm, ve = scene.poses(sample_envs, seed)

    def timed(threads, budget_s):
        """repeat passes over the sample until about budget_s of wall time has been spent"""
        orc.bench(xp, xm, ve, use_bvh=True, with_sensors=with_sensors, threads=threads)  # warm caches
        t, c, n = 0.0, 0, 0
        while t < budget_s:
            dt, dc, _ = orc.bench(xp, xm, ve, use_bvh=True, with_sensors=with_sensors, threads=threads)
            t, c, n = t + dt, c + dc, n + sample_envs
        return {"env_steps_per_s": n / t, "pair_evals_per_s": c / t, "envs": n, "seconds": t, "threads": threads}

    out = {"single_thread": timed(1, 6.0)}
    if threads_all:
        out["all_threads"] = timed(n_threads, 6.0)
    return out, n_threads


def run_reference(args, scene, with_sensors):
    """--impl reference: the reference's CPU implementation of the path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mujoco_contact_surfaces_b200 import scenes as S
    from oracle import oracle
    orc = oracle.OracleScene(scene.triangle, scene.apply_forces)
    S.configure(orc, scene)
    # all the host threads this process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently turn this arm into the one-thread variant for N > 1
    threads = max(oracle.num_threads(), len(os.sched_getaffinity(0)))
    xp, xm, ve = scene.poses(8, 1234)
    t8, _, _ = orc.bench(xp, xm, ve, True, with_sensors, 1)
    per_env = max(t8 / 8, 1e-7)
    budget = 120.0 / max(1, args.steps + args.warmup)  # whole run within a few minutes
    sample = int(min(args.envs, max(threads * 4, budget * threads / per_env)))
    sets = [scene.poses(sample, 1234 + i) for i in range(min(args.pose_sets, 4))]
    for i in range(args.warmup):
        orc.bench(*sets[i % len(sets)], True, with_sensors, threads)
    t, cands = 0.0, 0
    for i in range(args.steps):
        dt, c, _ = orc.bench(*sets[i % len(sets)], True, with_sensors, threads)
        t += dt
        cands += c
    val = sample * args.steps / t
    line = {
        "impl": "reference", "metric": "contact_surface_env_steps_per_sec", "value": val, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": scene.name, "envs_per_step_sample": sample, "note":
                   "reference CPU path restated (oracle port: Drake v1.8.0 + plugin force law + flat sensor); "
                   "the reference itself needs Drake/MuJoCo/ROS and cannot be built here"},
        "pair_evals_per_sec": cands / t,
        "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d envs per step, OpenMP over envs on all host threads" % sample},
        "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_flat(mesh, resolution, sampling, seconds):
    """CPU leg of benchmark_flat.py (the reference's benchmark_flat.cpp grid): ONE environment through the oracle,
    contact surface + flat-sensor update, use_parallel off / on (OpenMP over the taxels, flat_tactile_sensor.cpp:316)."""
    from mujoco_contact_surfaces_b200 import scenes as S
    from oracle import oracle
    scene = S.myrmex(mesh, sampling_resolution=sampling, resolution=resolution)
    orc = oracle.OracleScene(True, scene.apply_forces)
    S.configure(orc, scene)
    xp, xm, ve = scene.poses(4, seed=7)
    for mode, parallel in (("cpu_serial", False), ("cpu_parallel", True)):
        t, n, ts = 0.0, 0, 0.0
        while t < seconds and n < 100:
            e = n % 4
            t0 = time.perf_counter()
            orc.step(xp[e], xm[e], ve[e])
            t1 = time.perf_counter()
            orc.sensor_image(0, use_bvh=True, parallel=parallel)
            t2 = time.perf_counter()
            t, ts, n = t + (t2 - t0), ts + (t2 - t1), n + 1
        print(json.dumps(dict(impl=mode, n_envs=1, resolution=resolution, sampling_resolution=sampling, mesh=mesh,
                              surface_ms=1e3 * (t - ts) / n, sensor_ms=1e3 * ts / n, total_ms=1e3 * t / n,
                              ms_per_env=1e3 * t / n, nrays=0, ntri=len(orc.pair_triangles(0)))))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-flat":
        cpu_flat(sys.argv[2], float(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
        return
    args = parse()
    from mujoco_contact_surfaces_b200 import scenes as S
    scene = S.SCENES[args.workload]()
    with_sensors = bool(scene.sensors) if args.sensors < 0 else bool(args.sensors)
    if args.impl == "reference":
        run_reference(args, scene, with_sensors)
        return

    import torch
    import torch.distributed as dist
    from mujoco_contact_surfaces_b200 import REP_POLYGON, REP_TRIANGLE, HydroelasticEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the contact path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n_envs, ng, npairs = args.envs, scene.n_geoms, len(scene.pairs)
    # a dedicated non-default stream shared by torch and the engine, so torch.cuda.Event times our kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = HydroelasticEngine(n_envs, representation=REP_TRIANGLE if scene.triangle else REP_POLYGON,
                             apply_contact_forces=scene.apply_forces, device=local_rank, stream=stream.cuda_stream,
                             **scene.engine_kwargs(n_envs))
    S.configure(eng, scene)
    eng.finalize()

    # env shard of this rank: contiguous block [rank*n_envs, (rank+1)*n_envs); distinct pose sets per step
    sets_h, sets_d = [], []
    for i in range(args.pose_sets):
        xp, xm, ve = scene.poses(n_envs, seed=1234 + i, env_offset=rank * n_envs)
        hp = [torch.from_numpy(a.reshape(-1)).pin_memory() for a in (xp, xm, ve)]
        sets_h.append(hp)
        sets_d.append([t.to(dev) for t in hp])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_step(i):
        d = sets_d[i % len(sets_d)]
        eng.step_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), with_sensors)

    def host_step(i):
        h = sets_h[i % len(sets_h)]
        eng.step_raw(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), with_sensors)

    # ---------------- device-resident throughput (`value`) ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        dev_step(i)
    eng.sync()
    if rank == 0:
        t_wait = time.perf_counter()  # nvidia-smi needs a moment to start streaming: keep the GPU under load meanwhile
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:
            dev_step(0)
            eng.sync()
        sampler.rows.clear()
    stage = {"broadphase": 0.0, "narrowphase": 0.0, "reduce": 0.0, "tactile": 0.0, "setup": 0.0}
    cand = poly = faces = 0

    def timed_pass(with_stage_events):
        """K steps, each bracketed by its own pair of CUDA events on the engine's stream, L2 flushed in between.
        with_stage_events adds the engine's five per-stage event records inside every step: they give the kernel
        times the roofline needs, but each record costs the stream a few microseconds, so `value` comes from the
        pass WITHOUT them and the pass WITH them is reported next to it."""
        nonlocal cand, poly, faces
        eng.set_profiling(with_stage_events)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        w0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
            ev[i][0].record()
            dev_step(i)
            ev[i][1].record()
            eng.sync()  # also collects the per-stage CUDA events of this step
            if with_stage_events:
                sm = eng.stage_ms()
                for k in stage:
                    stage[k] += sm[k]
            else:
                c = eng.counters()  # D2H of the pair results, outside the events
                cand, poly, faces = cand + c["candidates"], poly + c["polygons"], faces + c["faces"]
        barrier()
        w1 = time.perf_counter()
        eng.set_profiling(False)
        return sum(a.elapsed_time(b) for a, b in ev), w1 - w0

    dev_ms, wall_s = timed_pass(False)            # the timed region of `value`
    wall0, wall1 = 0.0, wall_s
    if args.no_stage_events:
        staged_ms = dev_ms
    else:
        staged_ms, _ = timed_pass(True)           # same K steps again with the per-stage events (roofline kernel time)
    kernels = eng.counters()["kernels"]
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end to end through the C ABI with host buffers (`e2e`) ----------------
    for i in range(args.warmup):
        host_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        host_step(i)  # H2D of poses + kernels + D2H of the per-geom wrenches (/ images), synchronous
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = n_envs * ng * 18 * 8
    d2h = n_envs * ng * 48 + 16  # per-geom wrenches + flags (per-pair diagnostics stay on the device unless asked for)
    if with_sensors:
        d2h += sum(n_envs * cx * cy * 4 for cx, cy in eng.sensors)

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([cand, poly], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks (timing only; nothing on the step path)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    cand_all, poly_all = float(tot[0]), float(tot[1])

    if rank == 0:
        total_envs = n_envs * world
        value = total_envs * args.steps / (dev_ms_max * 1e-3)
        e2e_val = total_envs * args.steps / (e2e_ms_max * 1e-3)
        kinds = workload_kinds(scene)
        # Algorithmic bytes (DESIGN.md "Roofline"), rank 0, last step.  The dominant kernel is the
        # narrowphase: its launches process the pairs that survived the broadphase early-outs (soft-rigid,
        # soft-soft) or the tets the plane cuts (half space); each unit moves 232 / 264 / 132 B + 80 B per
        # emitted polygon.  The whole pair-eval pipeline (LBVH leaf hits / classified tets decided by broadphase +
        # narrowphase) is reported next to it against the time of both kernels.
        res = eng.pair_results()
        alg = alg_pipeline = 0.0
        for p, kind in enumerate(kinds):
            if kind:
                units = res["n_clipped"][:, p]
                alg += float(units.sum()) * BYTES_PER_PAIR[kind] + float(res["n_polygons"][:, p].sum()) * BYTES_PER_POLYGON
                alg_pipeline += float(res["n_candidates"][:, p].sum()) * BYTES_PER_PAIR[kind] + float(res["n_polygons"][:, p].sum()) * BYTES_PER_POLYGON
        n_narrow = sum(1 for k in kinds if k)
        narrow_ms = stage["narrowphase"] / args.steps
        pipe_ms = (stage["narrowphase"] + stage["broadphase"]) / args.steps
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback"
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        achieved = alg / (narrow_ms * 1e-3) / 1e9 if narrow_ms > 0 else 0.0
        achieved_pipe = alg_pipeline / (pipe_ms * 1e-3) / 1e9 if pipe_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        line = {
            "metric": "contact_surface_env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": scene.name, "envs_per_gpu": n_envs, "geoms": ng, "pairs": npairs,
                       "representation": "kTriangle" if scene.triangle else "kPolygon", "sensors": len(scene.sensors) if with_sensors else 0,
                       "parallelism": "env-sharded x%d, no collective on the step path" % world,
                       "l2": "256 MiB flush between timed iterations; %d distinct pose sets" % args.pose_sets},
            "pair_evals_per_sec": cand_all / (dev_ms_max * 1e-3),
            "pair_evals_per_env_step": cand_all / (total_envs * args.steps),
            "polygons_per_env_step": poly_all / (total_envs * args.steps),
            "clipped_pairs_per_env_step_rank0": float(res["n_clipped"].sum()) / n_envs,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "ms_per_step_with_stage_events": staged_ms / args.steps,
            "wall_ms_per_step_incl_flush_and_readback": 1e3 * (wall1 - wall0) / args.steps,
            "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(kernels) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "narrowphase (%d launch%s per step)" % (n_narrow, "" if n_narrow == 1 else "es"),
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_step": alg, "kernel_ms_per_step": narrow_ms,
                         "note": "meshes are shared by all envs and L2 resident: the kernel is FP64/latency bound, "
                                 "DRAM traffic (ncu) is far below the algorithmic bytes",
                         "pair_eval_pipeline": {"achieved": achieved_pipe, "frac": achieved_pipe / peak,
                                                "algorithmic_bytes_per_step": alg_pipeline,
                                                "kernels_ms_per_step": pipe_ms, "units": "LBVH leaf hits (pair-evals)"}},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            cb, threads = cpu_baseline(scene, 1234, args.cpu_sample_envs, with_sensors)
            a = cb["all_threads"]
            line["cpu_baseline"] = {"value": a["env_steps_per_s"], "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": "%d env-steps of the same workload (seeded poses), OpenMP over envs, %.1f s; single "
                                              "thread: %.1f env-steps/s over %d env-steps" % (a["envs"], a["seconds"], cb["single_thread"]["env_steps_per_s"],
                                                                                               cb["single_thread"]["envs"]),
                                    "single_thread_value": cb["single_thread"]["env_steps_per_s"],
                                    "pair_evals_per_sec": a["pair_evals_per_s"]}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
