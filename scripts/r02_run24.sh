#!/bin/bash
# step counters in two sets (the finalize kernel zeroes the next step's) vs one memset per step
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
HCS_STEP_MEMSET=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
  run "c1-twosets-$rep" X=1 -- --no-extra-workloads
  run "c1-memset-$rep" HCS_STEP_MEMSET=1 -- --no-extra-workloads
done
run "c4-twosets" X=1 -- --workload c4_objects_on_plane --no-extra-workloads
run "c4-memset" HCS_STEP_MEMSET=1 -- --workload c4_objects_on_plane --no-extra-workloads
run "c1-1env-twosets" X=1 -- --envs 1 --steps 500 --no-extra-workloads
run "c1-1env-memset" HCS_STEP_MEMSET=1 -- --envs 1 --steps 500 --no-extra-workloads
