#!/bin/bash
# full parity suite, headline bench, kernel table, ncu full capture of the fused CGS kernel (after the plain run exited 0)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?" >> gpurun_out/bench_n1.err
timeout 300 python scripts/cgs_bench.py f32 2,4,6,8,12,16,24,25,32,48,49,56,64,75,88,100 > gpurun_out/cgs_bench_f32.log 2>&1
timeout 300 python scripts/cgs_bench.py bf16 4,8,16,25,50,100,200 > gpurun_out/cgs_bench_bf16.log 2>&1
grep torch gpurun_out/cgs_bench_f32.log gpurun_out/cgs_bench_bf16.log | cut -c1-330
CMD="python bench.py --steps 4 --warmup 1 --prefill random --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_fused.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cgs_update_project -c 5 -f -o gpurun_out/prof_fused_r01 $CMD > gpurun_out/ncu_fused.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print(round(d["value"], 3), d["e2e"] and round(d["e2e"]["value"], 3), d["recurrence_only"]["ms_per_step"], d.get("cpu_baseline", {}).get("value"), d["roofline"])
print({k: (v["achieved_gbs"], round(v["ms_total"], 1)) for k, v in d["kernels"].items()})
PY
