#!/bin/bash
T=r2c
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_tests.log)
tail -8 gpurun_out/${T}_tests.log
grep -n "pinned gates\|full size 2x\|cfg4 400k" gpurun_out/${T}_tests.log | cut -c1-420
timeout 100 python tools/dev_split.py 2>&1 | grep rows
B200SP_TC_NOSPLIT=1 timeout 100 python tools/dev_split.py 2>&1 | grep rows
timeout 600 python bench.py --steps 20 --warmup 5 --no-m32 --no-gpu-native --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
for r in d['roofline']['families']: print(r['kernel'], r['launches_per_step'], round(r['ms_per_step'],3), round(r['frac'],4))
"
