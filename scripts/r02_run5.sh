#!/bin/bash
. scripts/r02_common.sh
for g in 1 2 4 8; do
run c2box-g$g HCS_RASTER_GROUP=$g -- --workload c2_myrmex_box --envs 1024 --steps 100 --no-extra-workloads
run c2plate-g$g HCS_RASTER_GROUP=$g -- --workload c2_myrmex_plate --envs 1024 --steps 100 --no-extra-workloads
run c5-g$g HCS_RASTER_GROUP=$g -- --workload c5_grasp_box --envs 1024 --steps 8 --warmup 3 --no-extra-workloads
done
