#!/bin/bash
# multi-device context + batched adapter on 2 GPUs: parity tests, hcs_multi bench line, torchrun bench with strong-scaling legs
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
mujoco_contact_surfaces_b200/plugin/test_plugin | grep -e batched -e timing
python bench.py --multi-devices 0 --steps 200 --warmup 20 | tee gpurun_out/r02t_multi1.json | cut -c1-600
python bench.py --multi-devices 0,1 --steps 200 --warmup 20 | tee gpurun_out/r02t_multi2.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02t_bench_2gpu.json 2> gpurun_out/r02t_bench_2gpu.err
tail -c 400 gpurun_out/r02t_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02t_bench_2gpu.json').read().strip().splitlines()[-1])
print('2 GPUs: value %.2f M e2e %.2f M' % (d['value']/1e6, d['e2e']['value']/1e6))
for k, w in d.get('workloads', {}).items():
    print(' weak', k, 'value %.3f M e2e %.3f M' % (w['value']/1e6, w['e2e']['value']/1e6))
for k, w in d.get('strong_scaling', {}).items():
    print(' strong', k, w['envs_total'], 'envs: value %.3f M e2e %.3f M  %.4f ms' % (w['value']/1e6, w['e2e']['value']/1e6, w['ms_per_step']))
PY
