#!/bin/bash
timeout 600 python bench.py --steps 5 --warmup 3 --no-m32 --no-cpu-baseline 2>gpurun_out/tmp_bench.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['vs_gpu_native'], d['baseline_gpu_native'])"
grep -i "accumulategrad" gpurun_out/tmp_bench.err | head -3
