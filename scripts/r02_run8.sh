#!/bin/bash
# limbs by magic-constant additions; prism test and 8 CTAs/SM in the flat traversal
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "" prism ft8; do
  lib=""; [ -n "$v" ] && lib="HCS_LIB=$V/libhcs_b200.$v.so"
  run "c1-$v" $lib X=1 -- --no-extra-workloads
  run "c3-$v" $lib X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
  run "c5-$v" $lib X=1 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
done
