#!/bin/bash
# flat broadphase: slots per batch
. scripts/r02_common.sh
for sl in 0 8 16 32; do
  for w in "c1_sphere_on_box 4096" "c3_soft_soft 4096" "c4_objects_on_plane 4096"; do
    set -- $w
    run "$1-slots$sl" HCS_FT_SLOTS=$sl -- --workload $1 --envs $2 --steps 200 --warmup 5 --no-extra-workloads
  done
done
run "c5-slots32" HCS_FT_SLOTS=32 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
run "c5-slots16" HCS_FT_SLOTS=16 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
