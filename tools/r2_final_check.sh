#!/bin/bash
# final round-2 validation on one GPU: the whole GPU suite, smoke(), both bench arms, the unchanged reference model
T=${1:-r2x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --detail gpurun_out/${T}_detail.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
timeout 600 python bench.py --model reference --no-gpu-native --no-m32 --no-cpu-baseline > gpurun_out/${T}_bench_refmodel.json 2> gpurun_out/${T}_bench_refmodel.err
timeout 600 python bench.py --model reference --attach-tape --no-gpu-native --no-m32 --no-cpu-baseline > gpurun_out/${T}_bench_refmodel_tape.json 2> gpurun_out/${T}_bench_refmodel_tape.err
python - <<PY
import json
for f in ("bench", "ref", "bench_refmodel", "bench_refmodel_tape"):
    try:
        d = json.loads(open("gpurun_out/${T}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("vs_gpu_native"), d.get("clocks"))
        if f == "bench":
            print(" roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic")})
            print(" m32", d.get("m32"), " cpu", d.get("cpu_baseline"), " parity", d.get("parity_full_size"), "launches", d.get("gpu_launches"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
