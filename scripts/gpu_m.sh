#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "graph" > gpurun_out/pytest_graph.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_graph.log | cut -c1-400
timeout 300 python bench.py --small --steps 10 --warmup 1 --no-extras > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "small exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_small.json')); print('small', d['value'], d['hvp_mode'][:60], d['e2e']['value'])"
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_n1_pipe.json 2> gpurun_out/bench_n1_pipe.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1_pipe.json"))
print(round(d["value"], 3), d["e2e"]["value"], d["ms_per_step"], d["hvp_ms_per_step"], d["recurrence_only"]["ms_per_step"], d["phases_ms_per_step"], d["ritz_top3"], d["hvp_mode"][-60:])
print({k: v["achieved_gbs"] for k, v in d["kernels"].items()})
PY
