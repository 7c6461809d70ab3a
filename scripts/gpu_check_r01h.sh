# validation of the final tree: GPU tests, smoke, C5 small batches (fine slices + cooperative finalize), C1 sanity
run() { # workload envs steps
  timeout 300 python bench.py --no-cpu-baseline --workload $1 --envs $2 --steps $3 --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$1', $2, round(d['value'],1), 'env-steps/s', round(d['ms_per_step'],4), 'ms | bp %.4f np %.4f red %.4f tac %.4f'%(s['broadphase'],s['narrowphase'],s['reduce'],s['tactile']), '| e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4))"
}
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -3
run c5_grasp_box 1 50
run c5_grasp_spot 1 50
run c5_grasp_box 4 50
run c5_grasp_box 16 30
run c1_sphere_on_box 4096 300
run c1_sphere_on_box 1 300
