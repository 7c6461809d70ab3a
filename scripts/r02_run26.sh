#!/bin/bash
# flat traversal: candidates flushed grouped by slot (long runs of equal environments for the narrowphase) vs as staged
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
  run "c1-sorted-$rep" X=1 -- --no-extra-workloads
  run "c1-unsorted-$rep" HCS_LIB=$V/libhcs_b200.unsorted.so -- --no-extra-workloads
done
run "c3-sorted" X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c3-unsorted" HCS_LIB=$V/libhcs_b200.unsorted.so -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c5-sorted" X=1 -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
run "c5-unsorted" HCS_LIB=$V/libhcs_b200.unsorted.so -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
