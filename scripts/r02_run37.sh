#!/bin/bash
# flat traversal: queries per batch between the powers of two (HCS_FT_SLOTS override): C1 has ~32 k alive queries for 2960 warps
. scripts/r02_common.sh
for sl in 0 16 14 12 11 10 8; do
  if [ $sl = 0 ]; then run "c1-default" X=1 -- --no-extra-workloads; else run "c1-slots$sl" HCS_FT_SLOTS=$sl -- --no-extra-workloads; fi
done
for sl in 0 12 10 6; do
  if [ $sl = 0 ]; then run "c4-default" X=1 -- --workload c4_objects_on_plane --steps 100 --no-extra-workloads; else run "c4-slots$sl" HCS_FT_SLOTS=$sl -- --workload c4_objects_on_plane --steps 100 --no-extra-workloads; fi
done
for sl in 0 24 20 12; do
  if [ $sl = 0 ]; then run "c3-default" X=1 -- --workload c3_soft_soft --steps 100 --no-extra-workloads; else run "c3-slots$sl" HCS_FT_SLOTS=$sl -- --workload c3_soft_soft --steps 100 --no-extra-workloads; fi
done
