# compute-sanitizer over the kernels added or changed late in round 1 (float leaf filters, face-vertex dump, area-importance
# taxel sampling, fine slices + cooperative finalize, C2b soft presser)
mkdir -p gpurun_out
K='float_leaf_filters and (unit or spheres_1mm) or face_vertices_are and (sphere_on_box-False or objects_on_plane-True or myrmex) or taxel_sensor_on_the_fingertip or taxel_sensor_on_the_myrmex_foam and area_importance or c5_grasp_full_resolution or c2_myrmex_taxel_image and soft_tip-8'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
