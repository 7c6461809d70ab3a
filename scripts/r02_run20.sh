#!/bin/bash
# C5 step kernels by time (256 envs), ncu launch list restricted to the step kernels
timeout 600 ncu -k regex:'bp_|broadphase_kernel|narrow_kernel|finalize|tactile_|scan_' --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 120 --csv --log-file gpurun_out/r02_launches_c5_grasp_box_256env.csv \
  python bench.py --workload c5_grasp_box --envs 256 --steps 3 --warmup 3 --no-cpu-baseline --no-extra-workloads --no-stage-events > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_launches_c5_grasp_box_256env.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr) and r[hdr.index('Metric Name')]=='gpu__time_duration.sum':
        k=r[hdr.index('Kernel Name')][:60]
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[hdr.index('Metric Value')].replace(',',''))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print('%-62s n=%3d  %10.1f us  %5.1f%%'%(k,v[0],v[1]/1e3,100*v[1]/tot))
PY
