#!/bin/bash
timeout 900 python -m pytest tests/test_reference_model_gpu.py -x -q -m gpu 2>&1 | tail -4
for f in "" "--attach-tape"; do
timeout 300 python bench.py --model reference $f --steps 30 --warmup 5 --no-gpu-native --no-m32 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('refmodel [$f]', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],2))"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-gpu-native --no-m32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mirror', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],2), [(f['kernel'][:14], round(f['ms_per_step'],2)) for f in d['roofline']['families']])"
