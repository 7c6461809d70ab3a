#!/bin/bash
# CUDA-graph hcs_step: parity, adapter timing, single-environment latency with and without the graph
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
make -C mujoco_contact_surfaces_b200/plugin -s
for g in "" "HCS_NO_GRAPH=1"; do
  echo "== adapter timing $g"
  env $g X=1 mujoco_contact_surfaces_b200/plugin/test_plugin | grep timing
done
for g in "" "HCS_NO_GRAPH=1"; do
for w in c1_sphere_on_box c4_objects_on_plane c2_myrmex_box c3_soft_soft; do
  env $g X=1 timeout 300 python bench.py --workload $w --envs 1 --steps 500 --warmup 20 --no-cpu-baseline --no-extra-workloads 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w $g 1 env: device %.1f us, hcs_step %.1f us, pipelined %.1f us' % (1e3*d['ms_per_step'], 1e3*d['e2e']['synchronous_hcs_step']['ms_per_step'], 1e3*d['e2e']['ms_per_step']))
    else: sys.stdout.write(l)
"
done
done
