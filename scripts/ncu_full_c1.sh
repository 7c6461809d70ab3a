# ncu --set full of the three C1 step kernels (one launch each, after warm-up), report brought back in gpurun_out/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'broadphase_kernel|narrow_tet_tri_kernel|finalize_env' \
  --launch-skip 9 -c 3 -f -o gpurun_out/r01_c1_final python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
