#!/bin/bash
T=r2f
mkdir -p gpurun_out
timeout 100 python tools/dev_split.py 2>&1 | grep rows
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_tests.log)
tail -4 gpurun_out/${T}_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d.get('vs_gpu_native'), d.get('m32'), d.get('parity_full_size'))
print(d['baseline_gpu_native'])
for r in d['roofline']['families']: print(r['kernel'], r['launches_per_step'], round(r['ms_per_step'],3), round(r['frac'],4))
"
