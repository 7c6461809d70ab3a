#!/bin/bash
# code size of the narrowphase kernels: tet-tet plane loop rolled (default) vs unrolled (ttunroll); accumulate column loop unrolled 16 (default) / 4 / 2
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
  for v in default ttunroll acc4 acc2; do
    if [ $v = default ]; then L=X=1; else L=HCS_LIB=$V/libhcs_b200.$v.so; fi
    run "c3-$v-$rep" $L -- --workload c3_soft_soft --steps 100 --no-extra-workloads
  done
done
for v in default acc4 acc2; do
  if [ $v = default ]; then L=X=1; else L=HCS_LIB=$V/libhcs_b200.$v.so; fi
  run "c1-$v" $L -- --no-extra-workloads
  run "c5-$v" $L -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
  run "c4-$v" $L -- --workload c4_objects_on_plane --steps 100 --no-extra-workloads
done
