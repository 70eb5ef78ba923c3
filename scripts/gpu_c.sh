#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python scripts/hvp_probe.py 1,2,4,8 > gpurun_out/hvp_probe.log 2>&1; echo "exit $?" >> gpurun_out/hvp_probe.log
HLV_TAG=stock timeout 300 python scripts/hvp_emul.py save > gpurun_out/hvp_emul_save.log 2>&1
PRE=/usr/local/cuda/lib64/libcublasLt.so.12:/usr/local/cuda/lib64/libcublas.so.12
HLV_TAG=preload_only LD_PRELOAD=$PRE timeout 300 python scripts/hvp_emul.py check > gpurun_out/hvp_emul_preload.log 2>&1
HLV_TAG=emul LD_PRELOAD=$PRE CUBLAS_EMULATE_SINGLE_PRECISION=1 timeout 300 python scripts/hvp_emul.py check > gpurun_out/hvp_emul_on.log 2>&1
HLV_TAG=emul_perf LD_PRELOAD=$PRE CUBLAS_EMULATE_SINGLE_PRECISION=1 CUBLAS_EMULATION_STRATEGY=performant timeout 300 python scripts/hvp_emul.py check > gpurun_out/hvp_emul_perf.log 2>&1
HLV_TAG=tf32 timeout 300 python - > gpurun_out/hvp_tf32.log 2>&1 <<'PY'
import torch, sys, os, json
sys.path.insert(0, '.')
import bench, hessian_llm_vision_b200 as hlv
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
model, cfg = bench.build_model(False); model.to(dev)
n = sum(p.numel() for p in model.parameters())
ids = bench.make_tokens(cfg, 8, 8, 512)[0].to(dev)
g = torch.Generator(device=dev).manual_seed(3)
v = torch.randn(n, device=dev, generator=g); v /= v.norm(); w = torch.empty(n, device=dev)
op = hlv.HessianVectorProduct(model, [ids])
for _ in range(2): op.accumulate_into(v, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): op.accumulate_into(v, w)
e1.record(); torch.cuda.synchronize()
ref = torch.load("gpurun_out/hv_ref.pt").to(dev)
print(json.dumps({"mode": "tf32", "ms": e0.elapsed_time(e1) / 5, "rel_err_vs_stock_fp32": float((w - ref).abs().max() / ref.abs().max()), "rel_l2_err": float((w - ref).norm() / ref.norm())}))
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err
rm -f gpurun_out/hv_ref.pt
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/hvp_probe.log | grep -v loss_type; tail -2 gpurun_out/hvp_emul_*.log gpurun_out/hvp_tf32.log; cat gpurun_out/bench_k20.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['kernels'])"
