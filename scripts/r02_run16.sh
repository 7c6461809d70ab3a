#!/bin/bash
# same-box sweep of launch-shape variants of the narrowphase / flat traversal (C1 and C5)
. scripts/r02_common.sh
for v in "" np3x7 np4x5 unroll2 xy128 ft5 np8x2 ""; do
  lib=""; [ -n "$v" ] && lib="HCS_LIB=$V/libhcs_b200.$v.so"
  run "c1-${v:-base}" $lib X=1 -- --no-extra-workloads
done
for v in "" np3x7 np4x5 ft5; do
  lib=""; [ -n "$v" ] && lib="HCS_LIB=$V/libhcs_b200.$v.so"
  run "c5-${v:-base}" $lib X=1 -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
  run "c3-${v:-base}" $lib X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
done
