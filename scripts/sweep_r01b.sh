# one-box A/B of the occupancy variants (built with HCS_VARIANT=..., see build.py); prints one line per run
run() { # name lib workload envs [env assignments...]
  n=$1; lib=$2; w=$3; envs=$4; shift 4
  env "$@" HCS_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --workload $w --envs $envs --steps 300 --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$n', '$w', round(d['value']/1e6,3), 'M', round(d['ms_per_step'],4), 'bp %.4f np %.4f red %.4f'%(s['broadphase'],s['narrowphase'],s['reduce']), 'e2e', round(d['e2e']['value']/1e6,3))"
}
D=$PWD/mujoco_contact_surfaces_b200/libhcs_b200.so
V=$PWD/mujoco_contact_surfaces_b200/variants
run default $D c1_sphere_on_box 4096 X=1
run tri5 $V/libhcs_b200.tri5.so c1_sphere_on_box 4096 X=1
run units8192 $D c1_sphere_on_box 4096 HCS_TARGET_UNITS=8192
run tri5_units8192 $V/libhcs_b200.tri5.so c1_sphere_on_box 4096 HCS_TARGET_UNITS=8192
run default $D c3_soft_soft 4096 X=1
run tet4 $V/libhcs_b200.tet4.so c3_soft_soft 4096 X=1
run default $D c4_objects_on_plane 4096 X=1
run plane5 $V/libhcs_b200.plane5.so c4_objects_on_plane 4096 X=1
run default $D c1_sphere_on_box 4096 X=1
run tri5 $V/libhcs_b200.tri5.so c1_sphere_on_box 4096 X=1
