#!/bin/bash
# final build on N GPUs: torchrun bench (weak headline + workloads + strong legs), hcs_multi from one process
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02s3_bench_${N}gpu.json 2> gpurun_out/r02s3_bench_${N}gpu.err
tail -c 300 gpurun_out/r02s3_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r02s3_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('$N GPUs: value %.2f M e2e %.2f M' % (d['value']/1e6, d['e2e']['value']/1e6))
for k, w in d.get('workloads', {}).items():
    print(' weak', k, 'value %.3f M e2e %.3f M' % (w['value']/1e6, w['e2e']['value']/1e6))
for k, w in d.get('strong_scaling', {}).items():
    print(' strong', k, w['envs_total'], 'envs: value %.3f M e2e %.3f M  %.4f ms' % (w['value']/1e6, w['e2e']['value']/1e6, w['ms_per_step']))
PY
DEV=$(python -c "print(','.join(str(i) for i in range($N)))")
timeout 300 python bench.py --multi-devices $DEV --steps 200 --warmup 20 | tee gpurun_out/r02s3_bench_hcs_multi_${N}gpu.json | cut -c1-330
