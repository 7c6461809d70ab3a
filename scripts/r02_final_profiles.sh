#!/bin/bash
# round-2 evidence run: tests, default bench + reference arm, ncu captures (full set of the C1 kernels and of the raster,
# launch lists of C1 / C2 / C5)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_final_tests.log; cat gpurun_out/r02_final_tests.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_final_bench_reference.json 2>/dev/null
timeout 1200 python bench.py --steps 200 --warmup 20 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
tail -c 300 gpurun_out/r02_final_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bp_prepare|bp_traverse|narrow_kernel|finalize' --launch-skip 16 -c 4 -f -o gpurun_out/r02_final_c1 \
  python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-extra-workloads > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tactile_raster_kernel' --launch-skip 6 -c 1 -f -o gpurun_out/r02_final_c2_raster \
  python bench.py --workload c2_myrmex_box --envs 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-extra-workloads > /dev/null 2>&1
for w in "c1_sphere_on_box 4096 c1" "c2_myrmex_box 1024 c2_myrmex_box" "c5_grasp_box 256 c5_grasp_box_256env"; do
  set -- $w
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$3.csv \
    python bench.py --workload $1 --envs $2 --steps 2 --warmup 1 --no-cpu-baseline --no-extra-workloads --no-stage-events > /dev/null 2>&1
done
ls -la gpurun_out/r02_final* gpurun_out/r02_launches*
