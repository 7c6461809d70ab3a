#!/bin/bash
# narrowphase: slot-list polygons, compact (non-unrolled) group loops vs r2base
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
  run "c1-new-$rep" X=1 -- --no-extra-workloads
  run "c1-base-$rep" HCS_LIB=$V/libhcs_b200.r2base.so -- --no-extra-workloads
done
run "c3-new" X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c3-base" HCS_LIB=$V/libhcs_b200.r2base.so -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c5-new" X=1 -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
run "c5-base" HCS_LIB=$V/libhcs_b200.r2base.so -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
