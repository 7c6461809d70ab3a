#!/bin/bash
source scripts/r02_common.sh
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
run c1
run c1-units8k HCS_TARGET_UNITS=8192
run c3 -- --workload c3_soft_soft --envs 4096 --steps 100
run c4 -- --workload c4_objects_on_plane --envs 4096 --steps 200
run c5 -- --workload c5_grasp_box --envs 1024 --steps 10 --warmup 3
