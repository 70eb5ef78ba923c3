#!/bin/bash
# fused (TMA slab) CGS middle pass: targeted tests under a short timeout first (a hung mbarrier must not eat the box)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "update_project" --timeout 120 > gpurun_out/pytest_fused.log 2>&1
rc=$?; echo "fused tests exit $rc" >> gpurun_out/pytest_fused.log
tail -15 gpurun_out/pytest_fused.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fused_k20.json 2> gpurun_out/bench_fused_k20.err; echo "exit $?" >> gpurun_out/bench_fused_k20.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-fused > gpurun_out/bench_unfused_k20.json 2> gpurun_out/bench_unfused_k20.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --basis-dtype bf16 > gpurun_out/bench_fused_bf16_k20.json 2> gpurun_out/bench_fused_bf16_k20.err
python - <<'PY'
import json
for f in ("bench_fused_k20", "bench_unfused_k20", "bench_fused_bf16_k20"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 3), d["recurrence_only"]["ms_per_step"], {k: (v["achieved_gbs"], round(v["ms_total"], 1)) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
