#!/bin/bash
# ncu --set full of the C1 step kernels (one launch each, after warm-up); report comes back in gpurun_out/
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bp_prepare|bp_traverse|broadphase_kernel|narrow_kernel|finalize_kernel' \
  --launch-skip 16 -c 4 -f -o gpurun_out/${1:-r02_c1} python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-extra-workloads > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
