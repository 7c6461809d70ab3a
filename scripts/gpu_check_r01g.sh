# final round-1 measurement of the current tree: smoke, headline bench + reference arm, launch list, other configs, one-env latencies
mkdir -p gpurun_out
( time python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c1.log 2>&1; tail -1 gpurun_out/bench_c1.log > gpurun_out/bench_c1.json; cut -c1-300 gpurun_out/bench_c1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c1_reference.json; cut -c1-200 gpurun_out/bench_c1_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
for w in c3_soft_soft c4_objects_on_plane; do timeout 300 python bench.py --workload $w --envs 4096 --steps 300 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_${w}_4096env.json; done
for w in c2_myrmex_box c2_myrmex_plate c2_myrmex_spot c3_soft_soft c4_objects_on_plane; do timeout 300 python bench.py --workload $w --envs 1024 --steps 200 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_${w}_1024env.json; done
timeout 600 python bench.py --workload c5_grasp_box --envs 1024 --steps 20 --warmup 3 --cpu-sample-envs 16 2>&1 | tail -1 > gpurun_out/bench_c5_grasp_box_1024env.json
for w in c1_sphere_on_box c2_myrmex_box c2_myrmex_spot c3_soft_soft c4_objects_on_plane c5_grasp_box; do timeout 300 python bench.py --workload $w --envs 1 --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${w}_1env.json; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*env.json')):
    try:
        d=json.load(open(f)); s=d['stage_ms_per_step']; cb=d.get('cpu_baseline') or {}
        print(f.split('/')[-1], '%.4g env-steps/s %.4f ms | bp %.4f np %.4f red %.4f tac %.4f | e2e %.4g (%.4f ms) | frac %.3f | cpu all %.4g 1t %.4g'%(d['value'],d['ms_per_step'],s['broadphase'],s['narrowphase'],s['reduce'],s['tactile'],d['e2e']['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],cb.get('value',0),cb.get('single_thread_value',0)))
    except Exception as e: print(f, 'ERR', e)
PY
