#!/bin/bash
# flat broadphase (prepare + traverse) vs the per-unit kernel: parity tests, then every workload both ways
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for leg in "" "HCS_BP_LEGACY=1"; do
  for w in "c1_sphere_on_box 4096" "c2_myrmex_box 1024" "c3_soft_soft 4096" "c4_objects_on_plane 4096" "c5_grasp_box 1024"; do
    set -- $w
    steps=200; [ "$1" = "c5_grasp_box" ] && steps=8
    run "$1${leg:+-legacy}" $leg X=1 -- --workload $1 --envs $2 --steps $steps --warmup 5 --no-extra-workloads
  done
done
