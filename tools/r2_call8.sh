#!/bin/bash
T=r2h
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_tests.log)
tail -4 gpurun_out/${T}_tests.log
for pdl in 1 0; do
B200SP_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-m32 --no-gpu-native --no-cpu-baseline --no-roofline > gpurun_out/${T}_bench_pdl$pdl.json 2> gpurun_out/${T}_bench_pdl$pdl.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_pdl$pdl.json'))
print('PDL=$pdl', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['step_ms_rank0'][:6])
"
done
timeout 200 python tools/bench_refkernels.py > gpurun_out/${T}_refkernels.md 2> gpurun_out/${T}_refkernels.err
tail -32 gpurun_out/${T}_refkernels.md | cut -c1-200
tail -2 gpurun_out/${T}_refkernels.err
