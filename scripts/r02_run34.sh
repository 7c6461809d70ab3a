#!/bin/bash
# traversal: 64-byte nodes and alive records as 256-bit loads; prepare: context block written by two threads, 5 CTAs / SM (96 registers); variants prep4 (112), prep6 (80, spills); vs r2base
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
  run "c1-new-$rep" X=1 -- --no-extra-workloads
  run "c1-prep4-$rep" HCS_LIB=$V/libhcs_b200.prep4.so -- --no-extra-workloads
  run "c1-prep6-$rep" HCS_LIB=$V/libhcs_b200.prep6.so -- --no-extra-workloads
  run "c1-base-$rep" HCS_LIB=$V/libhcs_b200.r2base.so -- --no-extra-workloads
done
for w in "c3_soft_soft --steps 100" "c4_objects_on_plane --steps 100" "c5_grasp_box --envs 512 --steps 8 --warmup 3" "c2_myrmex_spot --envs 1024 --steps 100"; do
  run "$w new" X=1 -- --workload $w --no-extra-workloads
  run "$w base" HCS_LIB=$V/libhcs_b200.r2base.so -- --workload $w --no-extra-workloads
done
