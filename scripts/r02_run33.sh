#!/bin/bash
# persistent kernels: a warp's first work item is its index in the grid (no atomic), later ones from the counter; vs r2base
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
  run "c1-new-$rep" X=1 -- --no-extra-workloads
  run "c1-base-$rep" HCS_LIB=$V/libhcs_b200.r2base.so -- --no-extra-workloads
done
one() { # label env-prefix workload envs
  env $2 timeout 300 python bench.py --workload $3 --envs $4 --steps 300 --warmup 20 --no-cpu-baseline --no-extra-workloads 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1 $3 $4 env: device %.1f us, hcs_step %.1f us, pipelined %.1f us' % (1e3*d['ms_per_step'], 1e3*d['e2e']['synchronous_hcs_step']['ms_per_step'], 1e3*d['e2e']['ms_per_step']), {k: round(1e3*v, 1) for k, v in d['stage_ms_per_step'].items()})
    elif 'rror' in l: sys.stdout.write(l)
"
}
for w in c1_sphere_on_box c4_objects_on_plane c3_soft_soft; do
  one new X=1 $w 1
  one base HCS_LIB=$V/libhcs_b200.r2base.so $w 1
done
for w in "c3_soft_soft --steps 100" "c4_objects_on_plane --steps 100" "c5_grasp_box --envs 512 --steps 8 --warmup 3"; do
  run "$w new" X=1 -- --workload $w --no-extra-workloads
  run "$w base" HCS_LIB=$V/libhcs_b200.r2base.so -- --workload $w --no-extra-workloads
done
