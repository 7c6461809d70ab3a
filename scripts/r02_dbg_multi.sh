#!/bin/bash
K="multi_device or multi or shard"
python -m pytest tests -m gpu -q -k "$K" --tb=short 2>&1 | grep -v "^$" | tail -12 | cut -c1-200
python -m pytest tests -m gpu -q -k "opt_in_survives" --tb=short 2>&1 | tail -5 | cut -c1-300
echo "== the regression test against the library before the fix"
HCS_LIB=mujoco_contact_surfaces_b200/variants/libhcs_b200.r2base.so python -m pytest tests -m gpu -q -k "opt_in_survives" --tb=line 2>&1 | tail -4 | cut -c1-300
python -m pytest tests -m gpu -q 2>&1 | tail -2
