#!/bin/bash
T=r2g
mkdir -p gpurun_out
timeout 600 python tools/microbench_cfg5.py > gpurun_out/${T}_cfg5.md 2> gpurun_out/${T}_cfg5.err
tail -30 gpurun_out/${T}_cfg5.md | cut -c1-220
timeout 400 python tools/bench_refkernels.py > gpurun_out/${T}_refkernels.md 2> gpurun_out/${T}_refkernels.err
tail -40 gpurun_out/${T}_refkernels.md | cut -c1-200
tail -3 gpurun_out/${T}_refkernels.err
