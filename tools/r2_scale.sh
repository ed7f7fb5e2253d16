#!/bin/bash
# multi-GPU bench runs: tools/r2_scale.sh N [extra bench flags]
N=$1; shift
T=r2s_n${N}$(echo "$*" | tr -d ' -')
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 --no-roofline "$@" > gpurun_out/${T}.json 2> gpurun_out/${T}.err < /dev/null
echo "N=$N flags [$*] rc $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], [round(v,1) for v in d["step_ms_rank0"]])
except Exception as e:
    print("no json", e); print(open("gpurun_out/${T}.err").read()[-1500:])
PY
