# one-box A/B: polygon vertices as (x, y) 128-bit + z 64-bit shared accesses (default) vs three 64-bit rows (variant polyrows)
run() { # name lib workload envs steps
  n=$1; lib=$2; w=$3; envs=$4; steps=$5
  HCS_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --workload $w --envs $envs --steps $steps --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$n', '$w', $envs, round(d['value']/1e6,4), 'M', round(d['ms_per_step'],4), 'bp %.4f np %.4f red %.4f tac %.4f'%(s['broadphase'],s['narrowphase'],s['reduce'],s['tactile']), 'e2e', round(d['e2e']['value']/1e6,4))"
}
D=$PWD/mujoco_contact_surfaces_b200/libhcs_b200.so
V=$PWD/mujoco_contact_surfaces_b200/variants/libhcs_b200.polyrows.so
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run xy128 $D c1_sphere_on_box 4096 300
run rows $V c1_sphere_on_box 4096 300
run xy128 $D c1_sphere_on_box 4096 300
run rows $V c1_sphere_on_box 4096 300
run xy128 $D c3_soft_soft 4096 100
run rows $V c3_soft_soft 4096 100
run xy128 $D c4_objects_on_plane 4096 300
run rows $V c4_objects_on_plane 4096 300
run xy128 $D c5_grasp_box 1024 20
run rows $V c5_grasp_box 1024 20
# fine slices for small batches (default) vs 32-query slices (HCS_NO_FINE_SLICES=1)
runs() { # workload envs steps
  for mode in fine coarse; do
    if [ $mode = coarse ]; then export HCS_NO_FINE_SLICES=1; else unset HCS_NO_FINE_SLICES; fi
    run $mode $D $1 $2 $3
  done
  unset HCS_NO_FINE_SLICES
}
runs c1_sphere_on_box 1 300
runs c1_sphere_on_box 16 300
runs c1_sphere_on_box 128 300
runs c1_sphere_on_box 1024 300
runs c1_sphere_on_box 2047 300
runs c3_soft_soft 1 200
runs c3_soft_soft 64 200
runs c2_myrmex_spot 1 200
runs c2_myrmex_spot 32 200
runs c5_grasp_box 1 50
runs c5_grasp_box 16 30
