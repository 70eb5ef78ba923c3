#!/bin/bash
# 2-GPU: NCCL invariance test + the headline line as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "multi_gpu" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_2gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_n2_k100.json 2> gpurun_out/bench_n2_k100.err
echo "exit $?"; wc -l gpurun_out/bench_n2_k100.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2_k100.json"))
print("N=2 value", round(d["value"], 3), d["hvp_mode"], "e2e", d["e2e"] and round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 2),
      "hvp", round(d["hvp_ms_per_step"], 2), "phases", d["phases_ms_per_step"], "ritz", d["ritz_top3"])
PY
