# one-box A/B: polygon tiles addressed by double index (default) vs by byte offset (variant bytesaddr); full GPU suite first
run() { # name lib workload envs steps
  n=$1; lib=$2; w=$3; envs=$4; steps=$5
  HCS_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --workload $w --envs $envs --steps $steps --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$n', '$w', $envs, round(d['value']/1e6,4), 'M', round(d['ms_per_step'],4), 'bp %.4f np %.4f red %.4f tac %.4f'%(s['broadphase'],s['narrowphase'],s['reduce'],s['tactile']), 'e2e', round(d['e2e']['value']/1e6,4))"
}
D=$PWD/mujoco_contact_surfaces_b200/libhcs_b200.so
V=$PWD/mujoco_contact_surfaces_b200/variants/libhcs_b200.bytesaddr.so
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run index $D c1_sphere_on_box 4096 300
run bytes $V c1_sphere_on_box 4096 300
run index $D c1_sphere_on_box 4096 300
run index $D c3_soft_soft 4096 100
run bytes $V c3_soft_soft 4096 100
run index $D c4_objects_on_plane 4096 200
run bytes $V c4_objects_on_plane 4096 200
run index $D c5_grasp_box 1024 15
run bytes $V c5_grasp_box 1024 15
