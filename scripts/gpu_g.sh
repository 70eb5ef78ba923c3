#!/bin/bash
# full parity suite, headline bench with the fused CGS pass, extras, kernel table, memcheck
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?" >> gpurun_out/bench_n1.err
timeout 900 python bench.py --extras --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_extras.json 2> gpurun_out/bench_n1_extras.err; echo "exit $?" >> gpurun_out/bench_n1_extras.err
timeout 300 python scripts/cgs_bench.py f32 4,8,10,16,25,26,50,51,75,100 > gpurun_out/cgs_bench_f32.log 2>&1
timeout 300 python scripts/cgs_bench.py bf16 8,25,50,100 > gpurun_out/cgs_bench_bf16.log 2>&1
timeout 120 python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
python - <<'PY'
import json
for f in ("bench_n1", "bench_n1_extras"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 3), d["e2e"] and round(d["e2e"]["value"], 3), d["recurrence_only"]["ms_per_step"], d.get("cpu_baseline", {}).get("value"), d["extras"], d["roofline"])
        print({k: (v["achieved_gbs"], round(v["ms_total"], 1)) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
grep torch gpurun_out/cgs_bench_f32.log | cut -c1-330; tail -5 gpurun_out/sanitize_memcheck.log
