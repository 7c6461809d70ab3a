#!/bin/bash
# subtree split of the flat traversal on large trees
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for sl in 0 4 16 32; do
  lib=""; [ $sl -gt 0 ] && lib="HCS_FT_SLOTS=$sl"
  run "c5-split-slots$sl" $lib X=1 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
done
run "c5-1env" X=1 -- --workload c5_grasp_box --envs 1 --steps 50 --warmup 5 --no-extra-workloads
run "c5-16env" X=1 -- --workload c5_grasp_box --envs 16 --steps 50 --warmup 5 --no-extra-workloads
run "c1" X=1 -- --no-extra-workloads
