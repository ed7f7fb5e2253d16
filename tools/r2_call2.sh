#!/bin/bash
T=r2b
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_tests.log)
tail -4 gpurun_out/${T}_tests.log
grep -n "full size\|cfg4\|reference model on\|mirror grads\|VGG net\|unet grad parity" gpurun_out/${T}_tests.log | cut -c1-600
timeout 100 python tools/dev_split.py > gpurun_out/${T}_split8.txt 2>&1
B200SP_TC_NOSPLIT=1 timeout 100 python tools/dev_split.py > gpurun_out/${T}_nosplit.txt 2>&1
B200SP_TC_MAXSPLIT=4 timeout 100 python tools/dev_split.py > gpurun_out/${T}_split4.txt 2>&1
B200SP_TC_MAXSPLIT=16 timeout 100 python tools/dev_split.py > gpurun_out/${T}_split16.txt 2>&1
tail -5 gpurun_out/${T}_split8.txt gpurun_out/${T}_nosplit.txt gpurun_out/${T}_split4.txt gpurun_out/${T}_split16.txt
timeout 600 python bench.py --steps 20 --warmup 5 --detail gpurun_out/${T}_detail.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.err
B200SP_LAYER_EXEC=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-gpu-native --no-cpu-baseline --no-m32 --no-roofline > gpurun_out/${T}_bench_nolayer.json 2> gpurun_out/${T}_bench_nolayer.err
timeout 300 python bench.py --model reference --steps 10 --warmup 3 --no-gpu-native --no-cpu-baseline > gpurun_out/${T}_bench_refmodel.json 2> gpurun_out/${T}_bench_refmodel.err
timeout 200 python tools/cpu_profile_step.py > gpurun_out/${T}_cpuprof.txt 2>&1
head -3 gpurun_out/${T}_cpuprof.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -c "
import json
for f in ('gpurun_out/${T}_bench.json','gpurun_out/${T}_bench_nolayer.json','gpurun_out/${T}_bench_refmodel.json'):
    d=json.load(open(f))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('vs_gpu_native'), d.get('parity_full_size',{}).get('scores_rel'), d.get('m32',{}).get('ms_per_step'), d['gpu_launches'])
d=json.load(open('gpurun_out/${T}_bench.json'))
for r in d['roofline']['families']: print(r['kernel'], r['launches_per_step'], round(r['ms_per_step'],3), round(r['frac'],4))
"
