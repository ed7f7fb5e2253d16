#!/bin/bash
T=r2e
mkdir -p gpurun_out
timeout 100 python tools/dev_split.py 2>&1 | grep rows
timeout 300 python tools/dev_wgrad_os.py big 2>&1 | grep "^rows"
timeout 300 python tools/dev_wgrad_os.py 2>&1 | grep "^rows\|strided\|inverse\|1x1"
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_tests.log)
tail -4 gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-m32 --no-gpu-native --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
for r in d['roofline']['families']: print(r['kernel'], r['launches_per_step'], round(r['ms_per_step'],3), round(r['frac'],4))
"
