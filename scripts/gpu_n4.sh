#!/bin/bash
# 8-GPU headline line exactly as the driver launches it (one rank per GPU, NCCL)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader > gpurun_out/gpus4.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 4 --steps 100 --warmup 3 > gpurun_out/bench_n4_k100.json 2> gpurun_out/bench_n4_k100.err
echo "exit $?" >> gpurun_out/bench_n4_k100.err
tail -3 gpurun_out/bench_n4_k100.err | cut -c1-400
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n4_k100.json"))
print("N=4 value", round(d["value"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 2),
      "hvp", round(d["hvp_ms_per_step"], 2), "phases", d["phases_ms_per_step"], "ritz", d["ritz_top3"])
PY
