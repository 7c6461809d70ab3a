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

The JSON line's headline is config 1 of BASELINE.json at its quoted size (soft sphere on rigid box x 4096 envs per
GPU).  The default invocation also measures, in short legs, the other configs (`workloads`: C2 Myrmex box S = 20,
C3 soft-soft, C4 mixed objects on a plane, C5 five 131 072-tet pads + taxel arrays) with value, e2e, dominant-kernel
roofline and a bounded cpu_baseline each, and under torchrun the STRONG-scaling legs BASELINE.json names (C4 with 4096
and C5 with 1024 environments IN TOTAL, split by env index: `strong_scaling`).  --no-extra-workloads leaves them out.
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
    ap.add_argument("--no-extra-workloads", action="store_true", help="only the headline workload (no `workloads` block)")
    ap.add_argument("--random-sizes", action="store_true", help="per-environment geom sizes (scenes that define them: C4)")
    ap.add_argument("--multi-devices", default="", help="e.g. 0,1: ONE process drives these GPUs through the library's "
                    "multi-device context (hcs_multi, C++ host threads); prints an end-to-end line, no torch.distributed")
    return ap.parse_args()


# short legs of the default invocation: (key, workload, envs per GPU (weak) , steps, warmup)
EXTRA_WORKLOADS = [("c2_myrmex_box_s20", "c2_myrmex_box", 1024, 60, 5), ("c3_soft_soft", "c3_soft_soft", 4096, 40, 5),
                   ("c4_objects_on_plane", "c4_objects_on_plane", 4096, 100, 10),
                   ("c4_objects_on_plane_random_sizes", "c4_objects_on_plane", 4096, 100, 10),  # per-environment sizes
                   ("c5_grasp_box", "c5_grasp_box", 1024, 6, 3)]
# strong-scaling legs under torchrun: total environments split over the ranks (BASELINE.json configs 4 and 5)
STRONG_WORKLOADS = [("c4_objects_on_plane", 4096, 100, 10), ("c5_grasp_box", 1024, 6, 3)]
# fp64 operations per pair-eval that reaches the clipper (SURVEY.md section 8d "Algorithmic flops"), polygon or not
FLOPS_PER_CLIPPED_PAIR = {"soft_rigid": 650.0, "soft_plane": 300.0, "soft_soft": 900.0}
FP64_PEAK_NO_FMA_TFLOPS = 18.5  # B200 vector FP64: 37 TFLOP/s counts an FMA as two; built with -fmad=false => half


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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def stop(self, t0=None, t1=None):
        """Samples between t0 and t1 (perf_counter): the window in which the GPU was kept under the measured load."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r[1:] for r in self.rows if len(r) >= 10 and (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1 + 0.05)]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "warm-up + timed passes (device-resident, stage-event and end-to-end), GPU under the measured load"}


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


class Rig:
    """What every measurement leg shares: the device, the torch stream the engine launches on, the L2 flush buffer, the
    rank layout."""

    def __init__(self, torch, dist, world, rank, local_rank):
        self.torch, self.dist, self.world, self.rank, self.local_rank = torch, dist, world, rank, local_rank
        self.dev = torch.device("cuda", local_rank)
        # a dedicated non-default stream shared by torch and the engine, so torch.cuda.Event times our kernels
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t]


def measure(rig, scene, n_envs, env_offset, steps, warmup, with_sensors, pose_sets, stage_events=True, sampler=None,
            random_sizes=False):
    """One leg: `steps` timed steps of `scene` on this rank's `n_envs` environments (global indices from env_offset).
    Returns this rank's raw numbers; the caller reduces them over the ranks."""
    torch = rig.torch
    from mujoco_contact_surfaces_b200 import REP_POLYGON, REP_TRIANGLE, HydroelasticEngine
    from mujoco_contact_surfaces_b200 import scenes as S
    ng = scene.n_geoms
    eng = HydroelasticEngine(n_envs, representation=REP_TRIANGLE if scene.triangle else REP_POLYGON,
                             apply_contact_forces=scene.apply_forces, device=rig.local_rank, stream=rig.stream.cuda_stream,
                             **scene.engine_kwargs(n_envs))
    S.configure(eng, scene)
    if random_sizes:  # domain-randomised geometry: per-environment meshes, fields and LBVHs (hcs_set_env_sizes)
        for g, sz in scene.env_sizes(n_envs, 1234, env_offset).items():
            eng.set_env_sizes(g, sz)
    t_fin = time.perf_counter()
    eng.finalize()
    finalize_s = time.perf_counter() - t_fin
    # env shard of this rank: contiguous block [env_offset, env_offset + n_envs); distinct pose sets per step
    sets_h, sets_d = [], []
    for i in range(pose_sets):
        xp, xm, ve = scene.poses(n_envs, seed=1234 + i, env_offset=env_offset)
        hp = [torch.from_numpy(a.reshape(-1)).pin_memory() for a in (xp, xm, ve)]
        sets_h.append(hp)
        sets_d.append([t.to(rig.dev) for t in hp])

    def dev_step(i):
        d = sets_d[i % len(sets_d)]
        eng.step_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), with_sensors)

    # ---------------- device-resident throughput (`value`) ----------------
    if sampler is not None:
        sampler.start()
    t_load0 = time.perf_counter()
    for i in range(warmup):
        dev_step(i)
    eng.sync()
    if sampler is not None:  # nvidia-smi needs a moment to start streaming: keep the GPU under the measured load meanwhile
        while not sampler.rows and time.perf_counter() - t_load0 < 3.0:
            dev_step(0)
            eng.sync()
    out = {"stage": {"broadphase": 0.0, "narrowphase": 0.0, "reduce": 0.0, "tactile": 0.0, "setup": 0.0}, "cand": 0, "poly": 0}

    def timed_pass(with_stage_events):
        """K steps, each bracketed by its own pair of CUDA events on the engine's stream, L2 flushed in between.
        with_stage_events adds the engine's five per-stage event records inside every step: they give the kernel
        times the roofline needs, but each record costs the stream a few microseconds, so `value` comes from the
        pass WITHOUT them and the pass WITH them is reported next to it."""
        eng.set_profiling(with_stage_events)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        rig.barrier()
        w0 = time.perf_counter()
        for i in range(steps):
            rig.flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
            ev[i][0].record()
            dev_step(i)
            ev[i][1].record()
            eng.sync()  # also collects the per-stage CUDA events of this step
            if with_stage_events:
                sm = eng.stage_ms()
                for k in out["stage"]:
                    out["stage"][k] += sm[k]
            else:
                c = eng.counters()  # D2H of the pair results, outside the events
                out["cand"] += c["candidates"]
                out["poly"] += c["polygons"]
        rig.barrier()
        w1 = time.perf_counter()
        eng.set_profiling(False)
        return sum(a.elapsed_time(b) for a, b in ev), w1 - w0

    out["dev_ms"], out["wall_s"] = timed_pass(False)  # the timed region of `value`
    out["staged_ms"] = timed_pass(True)[0] if stage_events else out["dev_ms"]
    out["kernels"] = eng.counters()["kernels"]
    out["finalize_s"] = finalize_s
    out["res"] = eng.pair_results()
    out["tactile_triangles"] = eng.counters()["tactile_triangles"]

    # ---------------- end to end through the C ABI with HOST buffers (`e2e`) ----------------
    # (a) hcs_step: synchronous (H2D of poses + kernels + D2H of the per-geom wrenches [+ images]): the latency of one call
    def host_step(i):
        h = sets_h[i % len(sets_h)]
        eng.step_raw(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), with_sensors)

    for i in range(warmup):
        host_step(i)
    rig.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        host_step(i)
    rig.barrier()
    out["e2e_sync_s"] = time.perf_counter() - t0
    # (b) hcs_step_async / hcs_wait: the same copies, two steps in flight (the H2D of step i+1 and the D2H of step i-1
    # overlap the kernels of step i); every step's wrenches (and taxel images) land in caller-owned pinned buffers and
    # the caller waits for step i-1 before it queues step i+1
    n_img = len(eng.sensors) if with_sensors else 0
    wr = [torch.empty(n_envs * ng * 6, dtype=torch.float64).pin_memory() for _ in range(2)]
    im = [[torch.empty(n_envs * cx * cy, dtype=torch.float32).pin_memory() for cx, cy in eng.sensors[:n_img]] for _ in range(2)]

    def pipe(n):
        prev = None
        for i in range(n):
            h = sets_h[i % len(sets_h)]
            t = eng.step_async(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), with_sensors, wr[i & 1].data_ptr(),
                               [b.data_ptr() for b in im[i & 1]] or None)
            if prev is not None:
                eng.wait(prev)
            prev = t
        if prev is not None:
            eng.wait(prev)

    pipe(warmup)
    rig.barrier()
    t0 = time.perf_counter()
    pipe(steps)
    rig.barrier()
    out["e2e_pipe_s"] = time.perf_counter() - t0
    out["wrench_checksum"] = float(wr[(steps - 1) & 1].abs().sum())
    out["h2d"] = n_envs * ng * 18 * 8
    out["d2h"] = n_envs * ng * 48 + 16  # per-geom wrenches + flags (per-pair diagnostics stay on the device unless asked for)
    if with_sensors:
        out["d2h"] += sum(n_envs * cx * cy * 4 for cx, cy in eng.sensors)
    out["n_sensors"] = len(eng.sensors) if with_sensors else 0
    out["sensor_cells"] = sum(cx * cy for cx, cy in eng.sensors) if with_sensors else 0
    out["rays_per_env"] = sum(cx * cy * int(sd["sampling_resolution"]) ** 2 for (cx, cy), sd in zip(eng.sensors, scene.sensors)) if with_sensors else 0
    out["clocks"] = sampler.stop(t_load0, time.perf_counter()) if sampler is not None else None
    eng.close()
    del sets_d
    return out


def peak_hbm():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def rooflines(scene, m, steps, n_envs, workload):
    """Roofline of the leg's dominant kernel (by stage time) + the narrowphase's, from rank 0's last step."""
    kinds = workload_kinds(scene)
    res = m["res"]
    alg = alg_pipeline = flops = 0.0
    for p, kind in enumerate(kinds):
        if kind:
            clipped, polys = float(res["n_clipped"][:, p].sum()), float(res["n_polygons"][:, p].sum())
            alg += clipped * BYTES_PER_PAIR[kind] + polys * BYTES_PER_POLYGON
            alg_pipeline += float(res["n_candidates"][:, p].sum()) * BYTES_PER_PAIR[kind] + polys * BYTES_PER_POLYGON
            flops += clipped * FLOPS_PER_CLIPPED_PAIR[kind]
    n_narrow = sum(1 for k in kinds if k)
    stage = {k: v / steps for k, v in m["stage"].items()}
    narrow_ms, pipe_ms = stage["narrowphase"], stage["narrowphase"] + stage["broadphase"]
    peak, peak_src = peak_hbm()
    achieved = alg / (narrow_ms * 1e-3) / 1e9 if narrow_ms > 0 else 0.0
    achieved_pipe = alg_pipeline / (pipe_ms * 1e-3) / 1e9 if pipe_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(workload)
    fp64 = flops / (narrow_ms * 1e-3) / 1e12 if narrow_ms > 0 else 0.0
    narrow = {"bound": "hbm", "kernel": "narrowphase (%d launch%s per step)" % (n_narrow, "" if n_narrow == 1 else "es"),
              "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
              "frac": achieved / peak, "traffic": traffic,
              "algorithmic_bytes_per_step": alg, "kernel_ms_per_step": narrow_ms,
              "note": "meshes are shared by all envs and L2 resident: the kernel is FP64/latency bound, "
                      "DRAM traffic (ncu) is far below the algorithmic bytes",
              "fp64": {"achieved": fp64, "peak": FP64_PEAK_NO_FMA_TFLOPS, "unit": "TFLOP/s", "frac": fp64 / FP64_PEAK_NO_FMA_TFLOPS,
                       "flops_per_clipped_pair": FLOPS_PER_CLIPPED_PAIR,
                       "note": "algorithmic fp64 operations of the clip + quadrature + force law per pair that reaches the "
                               "clipper (SURVEY.md section 8d), against the vector FP64 peak without FMA contraction "
                               "(-fmad=false is part of the parity contract)"},
              "pair_eval_pipeline": {"achieved": achieved_pipe, "frac": achieved_pipe / peak,
                                     "algorithmic_bytes_per_step": alg_pipeline,
                                     "kernels_ms_per_step": pipe_ms, "units": "LBVH leaf hits (pair-evals)"}}
    out = {"narrowphase": narrow, "dominant": narrow}
    if m["n_sensors"] and stage["tactile"] > max(stage["narrowphase"], stage["broadphase"]):
        # flat-sensor stage (bin + scan + raster): SURVEY.md section 8d figure = 48 B per (triangle, taxel) record + 4 B per
        # taxel; compute lens: one float32 Moeller-Trumbore evaluation (~40 flop) per sample ray of every taxel that holds
        # triangles is the floor (each lit sample needs at least its nearest triangle tested)
        tri = float(m["tactile_triangles"])
        cells = float(n_envs * m["sensor_cells"])
        rays = float(n_envs) * m["rays_per_env"]
        alg_t = 72.0 * tri + 4.0 * cells + 48.0 * tri  # pool records written + read once by the binning, images written
        t_ms = stage["tactile"]
        ach = alg_t / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        out["dominant"] = {"bound": "hbm", "kernel": "tactile stage (bin, scan, raster)", "achieved": ach, "peak": peak,
                           "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                           "algorithmic_bytes_per_step": alg_t, "kernel_ms_per_step": t_ms,
                           "fp32": {"achieved": rays * 40.0 / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0, "unit": "TFLOP/s",
                                    "peak": 74.0, "note": "40 flop per sample ray (one Moeller-Trumbore test per ray is the floor); "
                                                           "peak = 148 SMs x 128 FMA lanes x 2 x 1.965 GHz"},
                           "rays_per_step": rays, "tactile_triangles_per_step": tri}
    return out, stage


def leg_summary(rig, scene, m, steps, n_envs_rank, total_envs, workload, scaling):
    """Reduce one leg over the ranks (max of the times, sum of the counts) and shape it for the JSON line (rank 0)."""
    dev_ms, sync_ms, pipe_ms = rig.reduce([m["dev_ms"], m["e2e_sync_s"] * 1e3, m["e2e_pipe_s"] * 1e3], "MAX")
    cand, poly = rig.reduce([m["cand"], m["poly"]], "SUM")
    value = total_envs * steps / (dev_ms * 1e-3)
    roof, stage = rooflines(scene, m, steps, n_envs_rank, workload)
    return {
        "workload": scene.name, "envs_total": total_envs, "envs_this_rank": n_envs_rank, "scaling": scaling, "steps": steps,
        "value": value, "unit": "env-steps/s", "ms_per_step": dev_ms / steps,
        "pair_evals_per_sec": cand / (dev_ms * 1e-3), "pair_evals_per_env_step": cand / (total_envs * steps),
        "polygons_per_env_step": poly / (total_envs * steps),
        "clipped_pairs_per_env_step_rank0": float(m["res"]["n_clipped"].sum()) / max(1, n_envs_rank),
        "stage_ms_per_step": stage, "ms_per_step_with_stage_events": m["staged_ms"] / steps,
        "e2e": {"value": total_envs * steps / (pipe_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": m["h2d"],
                "d2h_bytes_per_step": m["d2h"], "ms_per_step": pipe_ms / steps,
                "call": "hcs_step_async + hcs_wait, pinned host buffers in and out, two steps in flight",
                "synchronous_hcs_step": {"value": total_envs * steps / (sync_ms * 1e-3), "ms_per_step": sync_ms / steps}},
        "gpu_launches": int(m["kernels"]) * steps,
        "roofline": roof["dominant"], "roofline_narrowphase": roof["narrowphase"],
        "wall_ms_per_step_incl_flush_and_readback": 1e3 * m["wall_s"] / steps,
    }


def cpu_leg(scene, with_sensors, budget_s):
    """Bounded cpu_baseline of an extra workload: the oracle on all host threads for about budget_s seconds."""
    from mujoco_contact_surfaces_b200 import scenes as S
    from oracle import oracle
    orc = oracle.OracleScene(scene.triangle, scene.apply_forces)
    S.configure(orc, scene)
    threads = max(oracle.num_threads(), len(os.sched_getaffinity(0)))
    xp, xm, ve = scene.poses(2, 1234)
    t2, _, _ = orc.bench(xp, xm, ve, True, with_sensors, 1)
    per_env = max(t2 / 2, 1e-7)
    sample = int(min(4096, max(threads, 0.5 * budget_s * threads / per_env)))
    xp, xm, ve = scene.poses(sample, 1234)
    t, n = 0.0, 0
    while t < budget_s and n < 50 * sample:
        dt, _, _ = orc.bench(xp, xm, ve, True, with_sensors, threads)
        t, n = t + dt, n + sample
    return {"value": n / t, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": "%d env-steps of the same workload, OpenMP over envs, %.1f s" % (n, t),
            "single_thread_value": 1.0 / per_env}


def run_multi(args, scene, with_sensors):
    """One process, one hcs_multi context over the listed devices: end-to-end env-steps/s through hcs_multi_step_async /
    hcs_multi_wait with pinned host buffers (weak scaling: args.envs environments per listed device)."""
    import torch
    from mujoco_contact_surfaces_b200 import REP_POLYGON, REP_TRIANGLE, MultiDeviceEngine
    from mujoco_contact_surfaces_b200 import scenes as S
    devices = [int(x) for x in args.multi_devices.split(",")]
    n_envs = args.envs * len(devices)
    eng = MultiDeviceEngine(n_envs, devices, representation=REP_TRIANGLE if scene.triangle else REP_POLYGON,
                            apply_contact_forces=scene.apply_forces, **scene.engine_kwargs(n_envs))
    S.configure(eng, scene)
    eng.finalize()
    ng = scene.n_geoms
    sets = []
    for i in range(min(args.pose_sets, 4)):
        xp, xm, ve = scene.poses(n_envs, seed=1234 + i)
        sets.append([torch.from_numpy(a.reshape(-1)).pin_memory() for a in (xp, xm, ve)])
    wr = [torch.empty(n_envs * ng * 6, dtype=torch.float64).pin_memory() for _ in range(2)]
    im = [[torch.empty(n_envs * cx * cy, dtype=torch.float32).pin_memory() for cx, cy in eng.sensors] if with_sensors else [] for _ in range(2)]

    def pipe(n):
        prev = None
        for i in range(n):
            h = sets[i % len(sets)]
            t = eng.step_async(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), with_sensors, wr[i & 1].data_ptr(),
                               [b.data_ptr() for b in im[i & 1]] or None)
            if prev is not None:
                eng.wait(prev)
            prev = t
        if prev is not None:
            eng.wait(prev)

    pipe(args.warmup)
    t0 = time.perf_counter()
    pipe(args.steps)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(args.steps):
        h = sets[i % len(sets)]
        eng._check(eng.L.hcs_multi_step(eng.h, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), int(with_sensors)))
    dt_sync = time.perf_counter() - t0
    print(json.dumps({"impl": "hcs_multi (one process, C++ host threads, no torch.distributed)", "devices": devices,
                      "metric": "contact_surface_env_steps_per_sec", "unit": "env-steps/s", "workload": scene.name,
                      "envs_total": n_envs, "steps": args.steps,
                      "e2e": {"value": n_envs * args.steps / dt, "ms_per_step": 1e3 * dt / args.steps,
                              "call": "hcs_multi_step_async + hcs_multi_wait, two steps in flight"},
                      "e2e_synchronous": {"value": n_envs * args.steps / dt_sync, "ms_per_step": 1e3 * dt_sync / args.steps},
                      "blocks": eng.blocks(), "wrench_checksum": float(wr[(args.steps - 1) & 1].abs().sum())}))
    eng.close()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-flat":
        cpu_flat(sys.argv[2], float(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
        return
    args = parse()
    from mujoco_contact_surfaces_b200 import scenes as S
    from mujoco_contact_surfaces_b200.sharding import shard_range
    scene = S.SCENES[args.workload]()
    with_sensors = bool(scene.sensors) if args.sensors < 0 else bool(args.sensors)
    if args.impl == "reference":
        run_reference(args, scene, with_sensors)
        return
    if args.multi_devices:
        run_multi(args, scene, with_sensors)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the contact path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rig = Rig(torch, dist, world, rank, local_rank)
    n_envs, ng, npairs = args.envs, scene.n_geoms, len(scene.pairs)

    # ---------------- headline: weak scaling, args.envs environments per GPU ----------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    m = measure(rig, scene, n_envs, rank * n_envs, args.steps, args.warmup, with_sensors, args.pose_sets,
                stage_events=not args.no_stage_events, sampler=sampler, random_sizes=args.random_sizes)
    head = leg_summary(rig, scene, m, args.steps, n_envs, n_envs * world, args.workload, "weak")

    # ---------------- the other configs (short legs) and the strong-scaling legs ----------------
    extra, strong = {}, {}
    if not args.no_extra_workloads and args.workload == "c1_sphere_on_box":
        for key, wl, envs, steps, warm in EXTRA_WORKLOADS:
            sc = S.SCENES[wl]()
            ws = bool(sc.sensors)
            rnd = key.endswith("_random_sizes")
            mm = measure(rig, sc, envs, rank * envs, steps, warm, ws, min(args.pose_sets, 4), random_sizes=rnd)
            leg = leg_summary(rig, sc, mm, steps, envs, envs * world, wl, "weak")
            if rnd:
                leg["geometry"] = ("per-environment sizes (hcs_set_env_sizes): every environment its own meshes, pressure fields and "
                                   "LBVHs, built on the GPU at finalize in %.2f s" % mm["finalize_s"])
            if rank == 0 and not args.no_cpu_baseline and not rnd:
                leg["cpu_baseline"] = cpu_leg(sc, ws, 2.5)
            extra[key] = leg
            del mm
        if world > 1:
            for wl, total, steps, warm in STRONG_WORKLOADS:
                sc = S.SCENES[wl]()
                ws = bool(sc.sensors)
                start, count = shard_range(total, rank, world)
                mm = measure(rig, sc, count, start, steps, warm, ws, min(args.pose_sets, 4))
                leg = leg_summary(rig, sc, mm, steps, count, total, wl, "strong")
                leg["note"] = ("%d environments IN TOTAL split by env index over %d GPUs (sharding.shard_range), no collective; "
                               "what limits the curve is the fixed per-step latency of the kernel chain at small shards" % (total, world))
                strong[wl] = leg
                del mm

    if rank == 0:
        line = {
            "metric": "contact_surface_env_steps_per_sec", "value": head["value"], "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": scene.name, "envs_per_gpu": n_envs, "geoms": ng, "pairs": npairs,
                       "representation": "kTriangle" if scene.triangle else "kPolygon", "sensors": len(scene.sensors) if with_sensors else 0,
                       "parallelism": "env-sharded x%d, no collective on the step path" % world,
                       "l2": "256 MiB flush between timed iterations; %d distinct pose sets" % args.pose_sets},
        }
        for k in ("pair_evals_per_sec", "pair_evals_per_env_step", "polygons_per_env_step", "clipped_pairs_per_env_step_rank0",
                  "stage_ms_per_step", "ms_per_step_with_stage_events", "wall_ms_per_step_incl_flush_and_readback", "e2e",
                  "gpu_launches", "roofline"):
            line[k] = head[k]
        line["clocks"] = m["clocks"]
        if not args.no_cpu_baseline:
            cb, threads = cpu_baseline(scene, 1234, args.cpu_sample_envs, with_sensors)
            a = cb["all_threads"]
            line["cpu_baseline"] = {"value": a["env_steps_per_s"], "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": "%d env-steps of the same workload (seeded poses), OpenMP over envs, %.1f s; single "
                                              "thread: %.1f env-steps/s over %d env-steps" % (a["envs"], a["seconds"], cb["single_thread"]["env_steps_per_s"],
                                                                                               cb["single_thread"]["envs"]),
                                    "single_thread_value": cb["single_thread"]["env_steps_per_s"],
                                    "pair_evals_per_sec": a["pair_evals_per_s"]}
        if extra:
            line["workloads"] = extra
        if strong:
            line["strong_scaling"] = strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
