#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --extras --no-cpu-baseline > gpurun_out/bench_n1_extras.json 2> gpurun_out/bench_n1_extras.err; echo "exit $?" >> gpurun_out/bench_n1_extras.err
timeout 600 python bench.py --steps 20 --warmup 3 --basis-dtype bf16 --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_bf16_k20.json 2> gpurun_out/bench_n1_bf16_k20.err; echo "exit $?" >> gpurun_out/bench_n1_bf16_k20.err
tail -15 gpurun_out/pytest_gpu.log | cut -c1-250; tail -2 gpurun_out/bench_n1_extras.err; python - <<'PY'
import json
for f in ("bench_n1_extras", "bench_n1_bf16_k20"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d.get("e2e", {}) and d["e2e"].get("value"), d["extras"], {k: (v["achieved_gbs"], v["frac_of_peak"]) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
