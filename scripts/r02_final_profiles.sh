#!/bin/bash
# round-2 evidence run: tests, default bench + reference arm, ncu captures (full set of the C1 kernels and of the raster,
# launch lists of C1 / C2 / C5)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02s3_final_tests.log; cat gpurun_out/r02s3_final_tests.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02s3_final_bench_reference.json 2>/dev/null
timeout 1200 python bench.py --steps 200 --warmup 20 > gpurun_out/r02s3_final_bench.json 2> gpurun_out/r02s3_final_bench.err
tail -c 300 gpurun_out/r02s3_final_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bp_prepare|bp_traverse|narrow_kernel|finalize' --launch-skip 16 -c 4 -f -o gpurun_out/r02s3_final_c1 \
  python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-extra-workloads > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tactile_raster_kernel' --launch-skip 6 -c 1 -f -o gpurun_out/r02s3_final_c2_raster \
  python bench.py --workload c2_myrmex_box --envs 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-extra-workloads > /dev/null 2>&1
# launch lists of the STEP kernels (the mesh / LBVH builders of hcs_finalize are filtered out by name)
for w in "c1_sphere_on_box 4096 c1" "c2_myrmex_box 1024 c2_myrmex_box" "c5_grasp_box 256 c5_grasp_box_256env"; do
  set -- $w
  timeout 600 ncu -k regex:'bp_|broadphase_kernel|narrow_kernel|finalize|tactile_|scan_' --metrics gpu__time_duration.sum --clock-control none --launch-skip 40 -c 160 --csv --log-file gpurun_out/r02s3_launches_$3.csv \
    python bench.py --workload $1 --envs $2 --steps 3 --warmup 3 --no-cpu-baseline --no-extra-workloads --no-stage-events > /dev/null 2>&1
done
# one environment through hcs_step and through the adapter
make -C mujoco_contact_surfaces_b200/plugin -s
mujoco_contact_surfaces_b200/plugin/test_plugin | grep -e timing -e batched > gpurun_out/r02s3_final_adapter.log
for w in c1_sphere_on_box c2_myrmex_box c3_soft_soft c4_objects_on_plane c5_grasp_box; do
  timeout 300 python bench.py --workload $w --envs 1 --steps 300 --warmup 20 --no-cpu-baseline --no-extra-workloads 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w 1 env: device %.1f us, hcs_step %.1f us, pipelined %.1f us' % (1e3*d['ms_per_step'], 1e3*d['e2e']['synchronous_hcs_step']['ms_per_step'], 1e3*d['e2e']['ms_per_step']))
" >> gpurun_out/r02s3_final_adapter.log
done
cat gpurun_out/r02s3_final_adapter.log
ls -la gpurun_out/r02s3_final* gpurun_out/r02_launches*
