#!/bin/bash
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "voxel" 2>&1 | tail -5
timeout 200 python tools/bench_augment.py 2>&1 | tail -6
