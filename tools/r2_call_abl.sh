#!/bin/bash
for cfg in "0 64" "1 200" "8 64" "8 110" "16 64" "0 64"; do set -- $cfg
B200SP_BN_CLUSTER=$1 B200SP_BN_CLUSTER_KB=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-gpu-native --no-m32 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['step_ms_rank0']; print('bn_cluster $cfg:', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],2), 'median', sorted(st)[len(st)//2])"
done
