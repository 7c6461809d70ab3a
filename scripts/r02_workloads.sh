#!/bin/bash
# all BASELINE.json configs on the current tree (device-resident + e2e), one summary line each
source scripts/r02_common.sh
run c1 -- --workload c1_sphere_on_box --envs 4096
run c2box -- --workload c2_myrmex_box --envs 1024 --steps 100
run c2spot -- --workload c2_myrmex_spot --envs 1024 --steps 100
run c3 -- --workload c3_soft_soft --envs 4096 --steps 100
run c4 -- --workload c4_objects_on_plane --envs 4096 --steps 200
run c5 -- --workload c5_grasp_box --envs 1024 --steps 10 --warmup 3
