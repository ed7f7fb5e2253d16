#!/bin/bash
# last GPU call of the round: suite, bench line, launch list of the final state, microbench sweep
set -x
T=${1:-r1f}
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python tools/microbench.py > gpurun_out/${T}_microbench.md 2> gpurun_out/${T}_microbench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${T}_ncu_bench.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'], d['cuda_mallocs_in_timed_steps'], d['step_ms_rank0'])
"
tail -3 gpurun_out/${T}_microbench.md
