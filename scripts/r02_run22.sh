#!/bin/bash
# one environment: pure kernel durations, warm caches (ncu --cache-control none) and cold
for cc in none all; do
for w in c1_sphere_on_box c4_objects_on_plane; do
  timeout 600 ncu -k regex:'bp_|broadphase_kernel|narrow_kernel|finalize' --metrics gpu__time_duration.sum --clock-control none --cache-control $cc --launch-skip 40 -c 24 --csv --log-file gpurun_out/r02_1env_$w.$cc.csv \
    python bench.py --workload $w --envs 1 --steps 30 --warmup 10 --no-cpu-baseline --no-extra-workloads --no-stage-events > /dev/null 2>&1
  python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_1env_$w.$cc.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr) and r[hdr.index('Metric Name')]=='gpu__time_duration.sum':
        k=r[hdr.index('Kernel Name')][:44]
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[hdr.index('Metric Value')].replace(',',''))
print('$w cache-control $cc:', ', '.join('%s x%d %.1f us' % (k, v[0], v[1]/v[0]/1e3) for k, v in agg.items()))
PY
done
done
