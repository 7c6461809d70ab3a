# one-box validation of the current tree: GPU parity tests, smoke, headline bench + reference arm, launch list, other configs
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
( time python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c1.log 2>&1; tail -1 gpurun_out/bench_c1.log > gpurun_out/bench_c1.json; cut -c1-600 gpurun_out/bench_c1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c1_reference.json; cut -c1-300 gpurun_out/bench_c1_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
bash scripts/bench_workloads.sh 4096 300 c3_soft_soft c4_objects_on_plane 2>&1 | tee gpurun_out/workloads.txt
bash scripts/bench_workloads.sh 1024 200 c2_myrmex_box c2_myrmex_plate c2_myrmex_spot 2>&1 | tee -a gpurun_out/workloads.txt
bash scripts/bench_workloads.sh 1024 20 c5_grasp_box 2>&1 | tee -a gpurun_out/workloads.txt
