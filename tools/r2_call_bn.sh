#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "bn" 2>&1 | tail -3
timeout 400 python tools/dev_bn.py 2>&1 | tail -62
for cfg in "0 64" "8 64" "8 200" "16 64" "8 64" "0 64"; do set -- $cfg
B200SP_BN_CLUSTER=$1 B200SP_BN_CLUSTER_KB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-gpu-native --no-m32 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', {k: d[k] for k in ('value','ms_per_step','gpu_launches')})"
done
