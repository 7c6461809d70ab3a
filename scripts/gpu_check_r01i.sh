# last measurement of round 1 (after the area-importance plumbing touched the kTriangle kernels): tests, headline, sensor configs
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_c1.log 2>&1; tail -1 gpurun_out/bench_c1.log > gpurun_out/bench_c1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c1_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
for w in c2_myrmex_box c2_myrmex_spot c2b_myrmex_soft_tip; do timeout 300 python bench.py --workload $w --envs 1024 --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${w}_1024env.json; done
timeout 600 python bench.py --workload c5_grasp_box --envs 1024 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c5_grasp_box_1024env.json
timeout 600 python bench.py --workload c5_grasp_box --envs 1 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c5_grasp_box_1env.json
python - <<'PY'
import json,glob
for f in ['gpurun_out/bench_c1.json']+sorted(glob.glob('gpurun_out/bench_c[25]*env.json')):
    try:
        d=json.load(open(f)); s=d['stage_ms_per_step']
        print(f.split('/')[-1], '%.4g env-steps/s %.4f ms | bp %.4f np %.4f red %.4f tac %.4f | e2e %.4g (%.4f ms) | frac %.3f'%(d['value'],d['ms_per_step'],s['broadphase'],s['narrowphase'],s['reduce'],s['tactile'],d['e2e']['value'],d['e2e']['ms_per_step'],d['roofline']['frac']))
    except Exception as e: print(f, 'ERR', e)
PY
