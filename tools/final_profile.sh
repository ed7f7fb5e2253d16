#!/bin/bash
# round-end evidence (one GPU): full bench line, ncu launch list of the bench command, ncu --set full of the kernels
# that carry the step.  Outputs under gpurun_out/ (copy the summaries into profiles/).
set -x
T=${1:-r1d}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${T}_ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
timeout 200 $NCU -k regex:k_conv_direct -o gpurun_out/${T}_full_conv_direct_l2 python tools/prof_one.py 118000 32 32 tc 3 > gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_conv_direct -o gpurun_out/${T}_full_conv_direct_l1 python tools/prof_one.py 300000 16 16 tc 3 >> gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_wgrad_direct -o gpurun_out/${T}_full_wgrad_direct_l1 python tools/prof_one.py 300000 16 16 tc 3 >> gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_wgrad_tc -o gpurun_out/${T}_full_wgrad_tc_l2 python tools/prof_one.py 118000 32 32 tc 3 >> gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_conv_tc -o gpurun_out/${T}_full_conv_tc_l3 python tools/prof_one.py 26500 48 48 tc 3 >> gpurun_out/${T}_full.log 2>&1
tail -3 gpurun_out/${T}_full.log
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d['cpu_baseline'], d['clocks'])
"
