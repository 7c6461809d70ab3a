#!/bin/bash
source scripts/r02_common.sh
for cfg in "c3 --workload c3_soft_soft --envs 4096 --steps 100" "c4 --workload c4_objects_on_plane --envs 4096 --steps 200" "c5 --workload c5_grasp_box --envs 1024 --steps 10 --warmup 3" "c1 --workload c1_sphere_on_box"; do
  set -- $cfg; name=$1; shift
  run $name-default -- "$@"
  run $name-upw1 HCS_BP_UPW=1 -- "$@"
  run $name-drain32 HCS_LIB=$V/libhcs_b200.drain32.so -- "$@"
  run $name-drain32-upw1 HCS_LIB=$V/libhcs_b200.drain32.so HCS_BP_UPW=1 -- "$@"
done
