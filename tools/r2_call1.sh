#!/bin/bash
# round 2, GPU call 1: new parity tests, bench (both models), host profile, smoke at several sizes
T=r2a
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_tests.log)
tail -5 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --detail gpurun_out/${T}_detail.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.err
timeout 300 python bench.py --model reference --steps 10 --warmup 3 --no-gpu-native --no-cpu-baseline > gpurun_out/${T}_bench_refmodel.json 2> gpurun_out/${T}_bench_refmodel.err
tail -c 300 gpurun_out/${T}_bench_refmodel.err
timeout 200 python tools/cpu_profile_step.py > gpurun_out/${T}_cpuprof.txt 2>&1
timeout 300 python tools/smoke_sizes.py > gpurun_out/${T}_smoke_sizes.txt 2>&1
cat gpurun_out/${T}_smoke_sizes.txt | tail -8
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('vs_gpu_native'), d.get('parity_full_size'), d.get('m32'))
print(d['roofline']['kernel'], d['roofline']['frac'])
"
