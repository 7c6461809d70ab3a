#!/bin/bash
# (tet, triangle) clip: first pass straight from the registers that hold the transformed triangle; vs the commit before (head)
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
  run "c1-new-$rep" X=1 -- --no-extra-workloads
  run "c1-head-$rep" HCS_LIB=$V/libhcs_b200.head.so -- --no-extra-workloads
done
run "c5 new" X=1 -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
run "c5 head" HCS_LIB=$V/libhcs_b200.head.so -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
run "c2 new" X=1 -- --workload c2_myrmex_spot --envs 1024 --steps 100 --no-extra-workloads
run "c2 head" HCS_LIB=$V/libhcs_b200.head.so -- --workload c2_myrmex_spot --envs 1024 --steps 100 --no-extra-workloads
