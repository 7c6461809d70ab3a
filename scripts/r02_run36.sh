#!/bin/bash
# finalize_env_kernel (<= 2 pairs): 128 / 64 / 32 environments per CTA (32 / 64 / 128 CTAs for 4096 environments)
. scripts/r02_common.sh
for rep in 1 2; do
  run "c1-fin128-$rep" X=1 -- --no-extra-workloads
  run "c1-fin64-$rep" HCS_LIB=$V/libhcs_b200.fin64.so -- --no-extra-workloads
  run "c1-fin32-$rep" HCS_LIB=$V/libhcs_b200.fin32.so -- --no-extra-workloads
done
run "c3 fin128" X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c3 fin32" HCS_LIB=$V/libhcs_b200.fin32.so -- --workload c3_soft_soft --steps 100 --no-extra-workloads
