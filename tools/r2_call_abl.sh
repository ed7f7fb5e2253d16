#!/bin/bash
for r in 0 1 2 3 4 5 6 7; do
timeout 300 python bench.py --as-rank $r --no-top-tape --steps 20 --warmup 5 --no-gpu-native --no-m32 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['step_ms_rank0']; print('scenes of rank $r alone:', round(d['ms_per_step'],2), 'ms/step', 'median', sorted(st)[len(st)//2], [ (l['rows']) for l in d['config']['levels']][:7])"
done
