#!/bin/bash
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run "c1" X=1 -- --no-extra-workloads
run "c3" X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads
run "c2box" X=1 -- --workload c2_myrmex_box --envs 1024 --steps 100 --no-extra-workloads
run "c2spot" X=1 -- --workload c2_myrmex_spot --envs 1024 --steps 100 --no-extra-workloads
run "c5" X=1 -- --workload c5_grasp_box --envs 512 --steps 8 --warmup 3 --no-extra-workloads
