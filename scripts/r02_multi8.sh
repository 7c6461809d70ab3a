#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device or refinalize" 2>&1 | tail -3
mujoco_contact_surfaces_b200/plugin/test_plugin | grep batched
for d in 0 0,1 0,1,2,3 0,1,2,3,4,5,6,7; do
  timeout 300 python bench.py --multi-devices $d --steps 300 --warmup 30 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(len(d['devices']), 'GPUs one process: e2e %.1f M (%.4f ms), synchronous %.1f M' % (d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e_synchronous']['value']/1e6))"
done
