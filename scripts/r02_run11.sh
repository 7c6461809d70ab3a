#!/bin/bash
. scripts/r02_common.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
run c4 -- --workload c4_objects_on_plane --no-extra-workloads
run c4-random-sizes -- --workload c4_objects_on_plane --random-sizes --no-extra-workloads
