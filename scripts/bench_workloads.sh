# usage: bash scripts/bench_workloads.sh <envs> <steps> workload...   (device-resident numbers only, no CPU baseline)
envs=$1; steps=$2; shift 2
for w in "$@"; do
  timeout 300 python bench.py --no-cpu-baseline --workload $w --envs $envs --steps $steps --warmup 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('$w', '$envs envs', round(d['value']), 'env-steps/s', round(d['ms_per_step'],4), 'ms | bp %.4f np %.4f red %.4f tac %.4f | e2e %d | pair-evals/s %.3g | roofline frac %.3f'%(s['broadphase'],s['narrowphase'],s['reduce'],s['tactile'],d['e2e']['value'],d['pair_evals_per_sec'],d['roofline']['frac']))"
done
