#!/bin/bash
# where does the time of the pooled broadphase / exact-accumulation narrowphase go (C1 x 4096 envs)
run() { # label, env assignments...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label', '%.2f M' % (d['value']/1e6), '%.4f ms' % d['ms_per_step'], {k: round(v, 4) for k, v in d['stage_ms_per_step'].items()}, 'clipped/env %.1f' % d['clipped_pairs_per_env_step_rank0'])
    else: sys.stdout.write(l)
"
}
V=mujoco_contact_surfaces_b200/variants
for u in 1 2 3 4 8; do run upw$u HCS_BP_UPW=$u; done
for v in noaccum noevals noprism noskip; do run $v HCS_LIB=$V/libhcs_b200.$v.so HCS_BP_UPW=2; done
