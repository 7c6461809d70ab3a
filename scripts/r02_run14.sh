#!/bin/bash
# per-kernel durations of the C4 step, flat vs per-unit half-space classification
for leg in "" "HCS_BP_LEGACY=1"; do
  env $leg X=1 timeout 600 ncu -k regex:'plane_classify|broadphase_kernel|narrow_kernel|finalize_kernel' --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --launch-skip 36 -c 9 --csv --log-file gpurun_out/r02_c4_launches${leg:+_legacy}.csv \
    python bench.py --workload c4_objects_on_plane --steps 3 --warmup 3 --no-cpu-baseline --no-extra-workloads > /dev/null 2>&1
  python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/r02_c4_launches${leg:+_legacy}.csv')))
hdr = None
cur = {}
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        key = (r[hdr.index('ID')], r[hdr.index('Kernel Name')][:48], r[hdr.index('Grid Size')])
        cur.setdefault(key, {})[r[hdr.index('Metric Name')][:24]] = r[hdr.index('Metric Value')]
for k, v in cur.items():
    print(k, v)
PY
done
