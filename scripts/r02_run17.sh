#!/bin/bash
# batch size of the flat traversal on large trees (C5) and C3
. scripts/r02_common.sh
for sl in 1 2 4 8 16; do
  run "c5-slots$sl" HCS_FT_SLOTS=$sl -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
done
for sl in 4 8 16; do
  run "c3-slots$sl" HCS_FT_SLOTS=$sl -- --workload c3_soft_soft --steps 100 --no-extra-workloads
  run "c2spot-slots$sl" HCS_FT_SLOTS=$sl -- --workload c2_myrmex_spot --envs 1024 --steps 100 --no-extra-workloads
done
run "c1" X=1 -- --no-extra-workloads
