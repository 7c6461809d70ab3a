#!/bin/bash
# A/B of the round-1 library against the current tree under ncu (instruction counts, issue utilisation) on C3 and C5
mkdir -p gpurun_out
SEC="--section SpeedOfLight --section LaunchStats --section Occupancy --section InstructionStats --section WarpStateStats --section SchedulerStats"
for lib in r1 cur; do
  if [ $lib = r1 ]; then export HCS_LIB=mujoco_contact_surfaces_b200/variants/libhcs_b200.r1.so; else unset HCS_LIB; fi
  timeout 600 ncu $SEC --clock-control none -k regex:'broadphase_kernel|narrow' --launch-skip 8 -c 2 -f -o gpurun_out/r02_ab_c3_$lib \
     python bench.py --workload c3_soft_soft --envs 4096 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ab.log 2>&1
  timeout 600 ncu $SEC --clock-control none -k regex:'broadphase_kernel' --launch-skip 20 -c 5 -f -o gpurun_out/r02_ab_c5_$lib \
     python bench.py --workload c5_grasp_box --envs 128 --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_ab.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
