#!/bin/bash
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
run c4 -- --workload c4_objects_on_plane --no-extra-workloads
run c1 -- --no-extra-workloads
run c1-1env -- --envs 1 --steps 500 --no-extra-workloads
run c5 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
