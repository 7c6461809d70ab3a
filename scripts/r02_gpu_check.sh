#!/bin/bash
# parity tests + C1 bench of the current tree (one B200); optional args are passed to bench.py
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('C1', '%.2f M env-steps/s' % (d['value']/1e6), '%.4f ms' % d['ms_per_step'], {k: round(v, 4) for k, v in d['stage_ms_per_step'].items()}, 'e2e %.2f M' % (d['e2e']['value']/1e6), 'clipped/env', d['clipped_pairs_per_env_step_rank0'], 'evals/env', d['pair_evals_per_env_step'], 'poly/env', d['polygons_per_env_step'])
    else: sys.stdout.write(l)
"
