#!/bin/bash
# narrowphase occupancy of the kPolygon (non-tactile) soft-rigid kernel: 4 warps x 5 CTAs (96 registers, 52 B spills) and 3 x 7 (80) vs 4 x 4 (123)
. scripts/r02_common.sh
for rep in 1 2; do
  run "c1-base-$rep" HCS_LIB=$V/libhcs_b200.r2base.so -- --no-extra-workloads
  run "c1-4x5-$rep" HCS_LIB=$V/libhcs_b200.base4x5.so -- --no-extra-workloads
  run "c1-3x7-$rep" HCS_LIB=$V/libhcs_b200.base3x7.so -- --no-extra-workloads
done
run "c4-base" HCS_LIB=$V/libhcs_b200.r2base.so -- --workload c4_objects_on_plane --steps 100 --no-extra-workloads
run "c4-4x5" HCS_LIB=$V/libhcs_b200.base4x5.so -- --workload c4_objects_on_plane --steps 100 --no-extra-workloads
run "c4-3x7" HCS_LIB=$V/libhcs_b200.base3x7.so -- --workload c4_objects_on_plane --steps 100 --no-extra-workloads
