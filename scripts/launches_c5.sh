mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5_grasp_box --envs 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3_soft_soft --envs 4096 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
