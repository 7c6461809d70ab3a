#!/bin/bash
# last single-GPU check of the round: compute-sanitizer over the step kernels, parity tests, default bench + reference arm
mkdir -p gpurun_out
# (compute-sanitizer: scripts/sanitize_r02.sh, run separately)
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02s3c_bench_reference.json 2>/dev/null
timeout 1200 python bench.py --steps 200 --warmup 20 > gpurun_out/r02s3c_bench.json 2> gpurun_out/r02s3c_bench.err
tail -c 300 gpurun_out/r02s3c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s3c_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r02s3c_bench_reference.json').read().strip().splitlines()[-1])
print('C1 value %.2f M  ms %.4f  e2e %.2f M sync %.2f M ref %.3f M ratio e2e %.1f'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['e2e']['synchronous_hcs_step']['value']/1e6,r['value']/1e6,d['e2e']['value']/r['value']))
print('stages',{k:round(v,4) for k,v in d['stage_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],4), 'clocks', d['clocks'])
for k,w in d.get('workloads',{}).items():
    print(k,'value %.4f M e2e %.4f M'%(w['value']/1e6,w['e2e']['value']/1e6),{a:round(b,4) for a,b in w['stage_ms_per_step'].items()},'frac',round(w['roofline']['frac'],3),'cpu',w.get('cpu_baseline',{}).get('value'))
PY
