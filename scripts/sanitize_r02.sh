#!/bin/bash
# compute-sanitizer over the kernels that are new in round 2 (flat broadphase, exact accumulation, finalize variants,
# raster with persistent warps, sphere generation, per-environment geometry, pipelined and multi-device steps)
mkdir -p gpurun_out
T="tests/test_gpu_parity.py"
SEL="c1_sphere_on_box_random or c3_soft_soft_polygon or c4_objects_on_plane or c2_myrmex_taxel_image or per_env_sizes or gpu_sphere_generation or pipelined_steps or multi_device or refinalize"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitize_$tool.log | tail -3
done
