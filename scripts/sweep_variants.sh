for n in "$@"; do
  HCS_LIB=$PWD/mujoco_contact_surfaces_b200/variants/libhcs_b200.$n.so timeout 200 python bench.py --no-cpu-baseline --steps 200 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$n', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],4), 'bp %.4f np %.4f red %.4f'%(s['broadphase'],s['narrowphase'],s['reduce']), 'e2e', round(d['e2e']['value']/1e6,2))"
done
