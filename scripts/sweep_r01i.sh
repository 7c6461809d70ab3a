# one-box A/B: float leaf filters (soft-rigid + soft-soft; default) vs the fp64-only leaf tests (variant noleaf32)
run() { # name lib workload envs steps
  n=$1; lib=$2; w=$3; envs=$4; steps=$5
  HCS_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --workload $w --envs $envs --steps $steps --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$n', '$w', $envs, round(d['value']/1e6,4), 'M', round(d['ms_per_step'],4), 'bp %.4f np %.4f red %.4f tac %.4f'%(s['broadphase'],s['narrowphase'],s['reduce'],s['tactile']), 'e2e', round(d['e2e']['value']/1e6,4), 'clipped/env', round(d['clipped_pairs_per_env_step_rank0'],2), 'evals/env', round(d['pair_evals_per_env_step'],2))"
}
D=$PWD/mujoco_contact_surfaces_b200/libhcs_b200.so
V=$PWD/mujoco_contact_surfaces_b200/variants/libhcs_b200.noleaf32.so
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run leaf32 $D c3_soft_soft 4096 100
run fp64 $V c3_soft_soft 4096 100
run leaf32 $D c3_soft_soft 4096 100
run fp64 $V c3_soft_soft 4096 100
run leaf32 $D c3_soft_soft 1024 100
run fp64 $V c3_soft_soft 1024 100
run leaf32 $D c1_sphere_on_box 4096 300
