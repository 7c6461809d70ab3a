#!/bin/bash
# prepare kernel: pose algebra once per environment and block (shared memory) vs once per thread
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
  run "c1-shared-$rep" X=1 -- --no-extra-workloads
  run "c1-perthread-$rep" HCS_LIB=$V/libhcs_b200.prepold.so -- --no-extra-workloads
done
run "c3-shared" X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c3-perthread" HCS_LIB=$V/libhcs_b200.prepold.so -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c2box-shared" X=1 -- --workload c2_myrmex_box --envs 1024 --steps 100 --no-extra-workloads
run "c2box-perthread" HCS_LIB=$V/libhcs_b200.prepold.so -- --workload c2_myrmex_box --envs 1024 --steps 100 --no-extra-workloads
