#!/bin/bash
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
make -C mujoco_contact_surfaces_b200/plugin -s
mujoco_contact_surfaces_b200/plugin/test_plugin | grep timing
for w in c1_sphere_on_box c4_objects_on_plane c2_myrmex_box; do
  timeout 300 python bench.py --workload $w --envs 1 --steps 300 --warmup 20 --no-cpu-baseline --no-extra-workloads 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w 1 env: device %.1f us, hcs_step %.1f us, pipelined %.1f us' % (1e3*d['ms_per_step'], 1e3*d['e2e']['synchronous_hcs_step']['ms_per_step'], 1e3*d['e2e']['ms_per_step']), {k: round(1e3*v, 1) for k, v in d['stage_ms_per_step'].items()})
    else: sys.stdout.write(l)
"
done
run c1 -- --no-extra-workloads
