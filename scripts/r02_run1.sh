#!/bin/bash
# GPU check 1 of round 2: parity tests, default bench (headline + workloads block), ncu of the tactile raster on C2 box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02k_tests.log
tail -3 gpurun_out/r02k_tests.log
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
tail -c 600 gpurun_out/r02k_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
    print('C1 value %.2f M e2e %.2f M (sync %.2f M) stages %s clocks %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['synchronous_hcs_step']['value']/1e6, {k: round(v, 4) for k, v in d['stage_ms_per_step'].items()}, d['clocks']))
    for k, w in d.get('workloads', {}).items():
        print(k, 'value %.3f M e2e %.3f M' % (w['value']/1e6, w['e2e']['value']/1e6), {a: round(b, 4) for a, b in w['stage_ms_per_step'].items()}, 'roof', w['roofline']['kernel'], round(w['roofline']['frac'], 3), 'cpu', w.get('cpu_baseline', {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tactile_raster_kernel' --launch-skip 6 -c 1 -f -o gpurun_out/r02k_c2_raster \
  python bench.py --workload c2_myrmex_box --envs 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02k_ncu.log 2>&1
tail -2 gpurun_out/r02k_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
