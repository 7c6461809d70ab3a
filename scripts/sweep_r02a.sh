# round-2 starter: A/B of the candidates read off the SASS at the end of round 1 (profiles/r01_notes.md); build first:
#   HCS_VARIANT=node256 HCS_NVCC_DEFS="-DHCS_BP_NODE256=1" python -m mujoco_contact_surfaces_b200.build
#   HCS_VARIANT=unroll2 HCS_NVCC_DEFS="-DHCS_CLIP_UNROLL=2" python -m mujoco_contact_surfaces_b200.build
run() { # name lib workload envs steps
  n=$1; lib=$2; w=$3; envs=$4; steps=$5
  HCS_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --workload $w --envs $envs --steps $steps --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$n', '$w', $envs, round(d['value']/1e6,4), 'M', round(d['ms_per_step'],4), 'bp %.4f np %.4f red %.4f tac %.4f'%(s['broadphase'],s['narrowphase'],s['reduce'],s['tactile']), 'e2e', round(d['e2e']['value']/1e6,4))"
}
D=$PWD/mujoco_contact_surfaces_b200/libhcs_b200.so
V=$PWD/mujoco_contact_surfaces_b200/variants
for v in node256 unroll2; do HCS_LIB=$V/libhcs_b200.$v.so timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1_ or c3_ or c5_grasp_five" 2>&1 | tail -2; done
for w in "c1_sphere_on_box 4096 300" "c3_soft_soft 4096 100" "c5_grasp_box 1024 15"; do
  run default $D $w
  for v in node256 unroll2; do run $v $V/libhcs_b200.$v.so $w; done
done
