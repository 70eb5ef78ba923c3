#!/bin/bash
# CUDA-graph HVP in the headline arms: functional check on the tiny model, then the N=1 line; graph-mode tests
mkdir -p gpurun_out
timeout 300 python bench.py --small --steps 10 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "small exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_small.json')); print('small', d['value'], d['hvp_mode'], d['e2e'])"
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "graph or adjust_implicit" > gpurun_out/pytest_graph.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_graph.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print(round(d["value"], 3), d["hvp_mode"], d["e2e"], d["recurrence_only"]["ms_per_step"], d["hvp_ms_per_step"], d["ritz_top3"], d["gpu_launches"])
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
