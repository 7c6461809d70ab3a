#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --multi-devices 0 --steps 300 --warmup 20 | cut -c1-420
python bench.py --multi-devices 0,0 --envs 2048 --steps 300 --warmup 20 | cut -c1-420
