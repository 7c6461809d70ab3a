#!/bin/bash
# final build on 2 GPUs: multi-device parity tests, torchrun bench (weak headline + workloads + strong legs), hcs_multi from one process
mkdir -p gpurun_out
nvidia-smi -L | head -2
python -m pytest tests -m gpu -q -k "multi_device or multi or shard" 2>&1 | tail -2
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02s3_bench_${N}gpu.json 2> gpurun_out/r02s3_bench_${N}gpu.err
tail -c 400 gpurun_out/r02s3_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r02s3_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('$N GPUs: value %.2f M e2e %.2f M' % (d['value']/1e6, d['e2e']['value']/1e6))
for k, w in d.get('workloads', {}).items():
    print(' weak', k, 'value %.3f M e2e %.3f M' % (w['value']/1e6, w['e2e']['value']/1e6))
for k, w in d.get('strong_scaling', {}).items():
    print(' strong', k, w['envs_total'], 'envs: value %.3f M e2e %.3f M  %.4f ms' % (w['value']/1e6, w['e2e']['value']/1e6, w['ms_per_step']))
PY
timeout 300 python bench.py --multi-devices 0,1 --steps 200 --warmup 20 | tee gpurun_out/r02s3_bench_hcs_multi_${N}gpu.json | cut -c1-400
