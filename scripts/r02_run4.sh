#!/bin/bash
# raster kernel with persistent warps: parity, C2 / C2b / C5 timings
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run c2box -- --workload c2_myrmex_box --envs 1024 --steps 100 --no-extra-workloads
run c2spot -- --workload c2_myrmex_spot --envs 1024 --steps 100 --no-extra-workloads
run c2plate -- --workload c2_myrmex_plate --envs 1024 --steps 100 --no-extra-workloads
run c2btip -- --workload c2b_myrmex_soft_tip --envs 1024 --steps 100 --no-extra-workloads
run c5 -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
