#!/bin/bash
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
run c1 -- --no-extra-workloads
run c5 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
