#!/bin/bash
timeout 200 python tools/host_sections.py 2000 2>&1 | grep -v Warn | tail -12
timeout 200 python tools/host_sections.py 150000 2>&1 | grep -v Warn | tail -12
for f in "" "--no-top-tape" "" "--no-top-tape"; do
timeout 300 python bench.py $f --steps 30 --warmup 5 --no-gpu-native --no-m32 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$f]', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],2))"
done
