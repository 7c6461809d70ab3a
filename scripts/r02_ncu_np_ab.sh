#!/bin/bash
# ncu --set full of the narrowphase kernel, new build vs r2base, on C3 and C5 (small batches)
V=mujoco_contact_surfaces_b200/variants
cap() { # name lib workload envs
  HCS_LIB=$2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'narrow_kernel' --launch-skip $5 -c 1 -f -o gpurun_out/$1 \
    python bench.py --workload $3 --envs $4 --steps 3 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-200
}
cap r02as_c3_new "" c3_soft_soft 1024 4
cap r02as_c3_base $V/libhcs_b200.r2base.so c3_soft_soft 1024 4
cap r02as_c5_new "" c5_grasp_box 32 20
cap r02as_c5_base $V/libhcs_b200.r2base.so c5_grasp_box 32 20
ls -la gpurun_out/r02as*
