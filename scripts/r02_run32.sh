#!/bin/bash
# fused small-step kernel (one thread-block cluster): parity tests, single-environment timings with and without it, small-batch sweep
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
make -C mujoco_contact_surfaces_b200/plugin -s
echo "== adapter timing fused"; mujoco_contact_surfaces_b200/plugin/test_plugin | grep timing
echo "== adapter timing HCS_NO_FUSED=1"; HCS_NO_FUSED=1 mujoco_contact_surfaces_b200/plugin/test_plugin | grep timing
one() { # label env-prefix workload envs
  env $2 timeout 300 python bench.py --workload $3 --envs $4 --steps 300 --warmup 20 --no-cpu-baseline --no-extra-workloads 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1 $3 $4 env: device %.1f us, hcs_step %.1f us, pipelined %.1f us, launches %s' % (1e3*d['ms_per_step'], 1e3*d['e2e']['synchronous_hcs_step']['ms_per_step'], 1e3*d['e2e']['ms_per_step'], d.get('gpu_launches')))
    elif 'rror' in l: sys.stdout.write(l)
"
}
for w in c1_sphere_on_box c4_objects_on_plane c3_soft_soft c2_myrmex_box; do
  one fused X=1 $w 1
  one chain HCS_NO_FUSED=1 $w 1
done
for n in 4 16 42 64 128; do
  one fused HCS_FUSED_MAX_QUERIES=1000000 c1_sphere_on_box $n
  one chain HCS_NO_FUSED=1 c1_sphere_on_box $n
done
for n in 2 4 8; do
  one fused HCS_FUSED_MAX_QUERIES=1000000 c4_objects_on_plane $n
  one chain HCS_NO_FUSED=1 c4_objects_on_plane $n
done
run c1 -- --no-extra-workloads
