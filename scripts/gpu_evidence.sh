#!/bin/bash
# round-1 evidence run (N=1): parity suite, smoke, both bench arms, eager-HVP and bf16 variants, ncu launch list of one step
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref exit $?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --hvp-mode eager --no-cpu-baseline > gpurun_out/bench_n1_eager.json 2> gpurun_out/bench_n1_eager.err; echo "eager exit $?"
timeout 900 python bench.py --basis-dtype bf16 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_bf16_k20.json 2> gpurun_out/bench_n1_bf16_k20.err; echo "bf16 exit $?"
CMD="python bench.py --steps 1 --warmup 0 --hvp-mode eager --prefill random --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_one_step.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import json
for f in ("bench_ref_n1", "bench_n1", "bench_n1_eager", "bench_n1_bf16_k20"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 3), d.get("hvp_mode"), d["e2e"] and round(d["e2e"]["value"], 3), d.get("recurrence_only", {}).get("ms_per_step"), d.get("roofline"))
    except Exception as e:
        print(f, "ERR", e)
PY
