#!/bin/bash
# per-env sizes tests; same-box A/B of the limb split (conversions vs additions), three alternations
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
for rep in 1 2 3; do
  run "c1-conv-$rep" X=1 -- --no-extra-workloads
  run "c1-add-$rep" HCS_LIB=$V/libhcs_b200.limbs.so -- --no-extra-workloads
done
HCS_LIB=$V/libhcs_b200.limbs.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
