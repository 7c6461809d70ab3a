#!/bin/bash
# flat traversal: sweep (every alive query against every tet box, 32 per pass) instead of the walk for trees up to 64 / 128 / 256 tets
. scripts/r02_common.sh
for rep in 1 2; do
  run "c1-walk-$rep" X=1 -- --no-extra-workloads
  run "c1-sweep128-$rep" HCS_LIB=$V/libhcs_b200.sweep128.so -- --no-extra-workloads
done
run "c1-sweep256" HCS_LIB=$V/libhcs_b200.sweep256.so -- --no-extra-workloads
for w in "c4_objects_on_plane --steps 100" "c2_myrmex_box --envs 1024 --steps 100" "c3_soft_soft --steps 100"; do
  run "$w walk" X=1 -- --workload $w --no-extra-workloads
  run "$w sweep128" HCS_LIB=$V/libhcs_b200.sweep128.so -- --workload $w --no-extra-workloads
  run "$w sweep256" HCS_LIB=$V/libhcs_b200.sweep256.so -- --workload $w --no-extra-workloads
done
