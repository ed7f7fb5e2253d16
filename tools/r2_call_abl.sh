#!/bin/bash
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_model_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -4
for r in 1 0 1 0; do
B200SP_TAPE_RAW=$r timeout 300 python bench.py --steps 30 --warmup 5 --no-gpu-native --no-m32 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('raw=$r', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],2), 'mallocs', d.get('cuda_mallocs_in_timed_steps'))"
done
for r in 1 0; do B200SP_TAPE_RAW=$r timeout 200 python tools/host_sections.py 2000 2>&1 | grep -v Warn | head -4; done
