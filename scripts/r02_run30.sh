#!/bin/bash
# rolled accumulate_chunk (conversion after the exchange) with the slot-list clip (new) and with the two-buffer clip (oldclip_newacc) vs r2base
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
HCS_LIB=$V/libhcs_b200.oldclip_newacc.so python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
  run "c1-new-$rep" X=1 -- --no-extra-workloads
  run "c1-oldclip_newacc-$rep" HCS_LIB=$V/libhcs_b200.oldclip_newacc.so -- --no-extra-workloads
  run "c1-base-$rep" HCS_LIB=$V/libhcs_b200.r2base.so -- --no-extra-workloads
done
for w in "c3_soft_soft --steps 100" "c4_objects_on_plane --steps 100" "c5_grasp_box --envs 512 --steps 8 --warmup 3"; do
  run "$w new" X=1 -- --workload $w --no-extra-workloads
  run "$w oldclip_newacc" HCS_LIB=$V/libhcs_b200.oldclip_newacc.so -- --workload $w --no-extra-workloads
  run "$w base" HCS_LIB=$V/libhcs_b200.r2base.so -- --workload $w --no-extra-workloads
done
