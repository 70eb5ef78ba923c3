#!/bin/bash
# what the driver runs at round end: parity suite, smoke, both bench arms (N=1)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; wc -l gpurun_out/bench_n1.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print(round(d["value"], 3), d["e2e"]["value"], d["hbm_roofline"], d["roofline"]["kernel"], d["roofline"]["frac"], d["extras"].keys())
PY
