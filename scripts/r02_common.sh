# helpers of the round-2 sweep scripts: `run label [ENV=value ...] [-- bench args]` prints one summary line of bench.py
V=mujoco_contact_surfaces_b200/variants
run() {
  local label=$1; shift
  local envs=() args=()
  while [ $# -gt 0 ]; do if [ "$1" = "--" ]; then shift; args=("$@"); break; fi; envs+=("$1"); shift; done
  env "${envs[@]}" timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline "${args[@]}" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label', '%.2f M' % (d['value']/1e6), '%.4f ms' % d['ms_per_step'], {k: round(v, 4) for k, v in d['stage_ms_per_step'].items()}, 'e2e %.2f M' % (d['e2e']['value']/1e6), 'clipped/env %.1f' % d['clipped_pairs_per_env_step_rank0'], 'evals/env %.1f' % d['pair_evals_per_env_step'])
    else: sys.stdout.write(l)
"
}
