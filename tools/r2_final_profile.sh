#!/bin/bash
# round-2 evidence (one GPU): GPU test suite, smoke, both bench arms, ncu launch list of the bench command, ncu --set full
# of the kernels that carry the step.  Outputs under gpurun_out/ (summaries are copied into profiles/ afterwards).
T=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --detail gpurun_out/${T}_detail.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
timeout 600 python bench.py --model reference --steps 20 --warmup 5 --no-gpu-native --no-m32 --no-cpu-baseline > gpurun_out/${T}_bench_refmodel.json 2> gpurun_out/${T}_bench_refmodel.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-gpu-native --no-m32 > gpurun_out/${T}_ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
timeout 200 $NCU -k regex:k_conv_tc -o gpurun_out/${T}_full_conv_tc_l3 python tools/prof_one.py 26500 48 48 tc 3 > gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_wgrad_os -o gpurun_out/${T}_full_wgrad_os_l2 python tools/prof_one.py 118000 32 32 tc 3 >> gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_wgrad_tc -o gpurun_out/${T}_full_wgrad_tc_l3 python tools/prof_one.py 26500 48 48 tc 3 >> gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_conv_direct -o gpurun_out/${T}_full_conv_direct_l2 python tools/prof_one.py 118000 32 32 tc 3 >> gpurun_out/${T}_full.log 2>&1
timeout 200 $NCU -k regex:k_bn_reduce -o gpurun_out/${T}_full_bn_reduce_l1 python tools/dev_bn.py run >> gpurun_out/${T}_full.log 2>&1
tail -3 gpurun_out/${T}_full.log
python - <<PY
import json
for f in ("bench", "ref", "bench_refmodel"):
    try:
        d = json.loads(open("gpurun_out/${T}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("vs_gpu_native"), d.get("clocks"))
        if f == "bench":
            print(" roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic")})
            print(" gpu_native", {k: v for k, v in (d.get("baseline_gpu_native") or {}).items() if k != "graph_error"}, (d.get("baseline_gpu_native") or {}).get("graph_error", "")[:300])
            print(" m32", d.get("m32"), " cpu", d.get("cpu_baseline"), " parity", d.get("parity_full_size"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
