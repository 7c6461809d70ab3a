#!/bin/bash
# narrowphase shared-memory traffic: duplicate removal only stores a vertex that moves, kPolygon only computes the pressure of vertex 0; vs the commit before (head)
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
  run "c1-new-$rep" X=1 -- --no-extra-workloads
  run "c1-head-$rep" HCS_LIB=$V/libhcs_b200.head.so -- --no-extra-workloads
done
for w in "c3_soft_soft --steps 100" "c4_objects_on_plane --steps 100" "c5_grasp_box --envs 512 --steps 8 --warmup 3"; do
  run "$w new" X=1 -- --workload $w --no-extra-workloads
  run "$w head" HCS_LIB=$V/libhcs_b200.head.so -- --workload $w --no-extra-workloads
done
