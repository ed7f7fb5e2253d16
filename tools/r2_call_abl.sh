#!/bin/bash
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/dev_split.py 2>&1 | grep rows
for v in 1 2; do
timeout 300 python bench.py --steps 40 --warmup 5 --no-gpu-native --no-m32 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')})"
done
