#!/bin/bash
# per-lane walk kernel vs the shared-queue traversal
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for leg in "" "HCS_BP_QUEUES=1"; do
  run "c1 $leg" $leg X=1 -- --no-extra-workloads
  run "c3 $leg" $leg X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
  run "c5 $leg" $leg X=1 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
  run "c2spot $leg" $leg X=1 -- --workload c2_myrmex_spot --envs 1024 --steps 100 --no-extra-workloads
done
