#!/bin/bash
# gathered geometry records (tet fields, tet vertices, triangles) loaded with L1::evict_last (default) vs plain loads (nokeep)
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
  run "c1-keep-$rep" X=1 -- --no-extra-workloads
  run "c1-nokeep-$rep" HCS_LIB=$V/libhcs_b200.nokeep.so -- --no-extra-workloads
done
for w in "c3_soft_soft --steps 100" "c4_objects_on_plane --steps 100" "c5_grasp_box --envs 512 --steps 8 --warmup 3" "c2_myrmex_box --envs 1024 --steps 100"; do
  run "$w keep" X=1 -- --workload $w --no-extra-workloads
  run "$w nokeep" HCS_LIB=$V/libhcs_b200.nokeep.so -- --workload $w --no-extra-workloads
done
