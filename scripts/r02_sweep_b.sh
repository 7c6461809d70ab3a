#!/bin/bash
source scripts/r02_common.sh
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
run default
run upw2 HCS_BP_UPW=2
run walk HCS_LIB=$V/libhcs_b200.walk.so
