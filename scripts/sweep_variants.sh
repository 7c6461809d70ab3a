# usage: bash scripts/sweep_variants.sh <workload> <envs> variant...   (variants built with HCS_VARIANT=..., see build.py)
w=$1; envs=$2; shift 2
for n in "$@"; do
  HCS_LIB=$PWD/mujoco_contact_surfaces_b200/variants/libhcs_b200.$n.so timeout 200 python bench.py --no-cpu-baseline --workload $w --envs $envs --steps 200 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$n', '$w', round(d['value']/1e6,3), 'M', round(d['ms_per_step'],4), 'bp %.4f np %.4f red %.4f'%(s['broadphase'],s['narrowphase'],s['reduce']), 'e2e', round(d['e2e']['value']/1e6,3))"
done
