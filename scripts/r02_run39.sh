#!/bin/bash
# read-once data (context blocks, candidate records, alive records) loaded with L1::no_allocate (default) vs plain loads (nostream)
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
  run "c1-stream-$rep" X=1 -- --no-extra-workloads
  run "c1-nostream-$rep" HCS_LIB=$V/libhcs_b200.nostream.so -- --no-extra-workloads
done
for w in "c3_soft_soft --steps 100" "c4_objects_on_plane --steps 100" "c5_grasp_box --envs 512 --steps 8 --warmup 3" "c2_myrmex_box --envs 1024 --steps 100"; do
  run "$w stream" X=1 -- --workload $w --no-extra-workloads
  run "$w nostream" HCS_LIB=$V/libhcs_b200.nostream.so -- --workload $w --no-extra-workloads
done
