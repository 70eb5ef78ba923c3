"""GPU probe (dev tool): does cuBLAS 12.9's BF16x9 fp32 emulation (tensor cores) reproduce the fp32 HVP?
  python scripts/hvp_emul.py save    # stock torch cuBLAS (SIMT sgemm): saves Hv + timing
  LD_PRELOAD=<cuda 12.9 libcublasLt:libcublas> CUBLAS_EMULATE_SINGLE_PRECISION=1 python scripts/hvp_emul.py check
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import hessian_llm_vision_b200 as hlv

mode = sys.argv[1]
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
model, cfg = bench.build_model(False)
model.to(dev)
n = sum(p.numel() for p in model.parameters())
ids = bench.make_tokens(cfg, 8, 8, 512)[0].to(dev)
g = torch.Generator(device=dev).manual_seed(3)
v = torch.randn(n, device=dev, generator=g); v /= v.norm()
w = torch.empty(n, device=dev)
op = hlv.HessianVectorProduct(model, [ids])
for _ in range(2):
    op.accumulate_into(v, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    op.accumulate_into(v, w)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
info = {"mode": mode, "ms": ms, "cublas_version": torch.backends.cuda.__dict__.get("cublas_version", None),
        "env": {k: os.environ.get(k) for k in ("LD_PRELOAD", "CUBLAS_EMULATE_SINGLE_PRECISION", "CUBLAS_EMULATION_STRATEGY")}}
if mode == "save":
    torch.save(w.cpu(), "gpurun_out/hv_ref.pt")
    # fp64 truth on a subset is too costly; self-consistency: second run bitwise?
    w2 = torch.empty_like(w); op.accumulate_into(v, w2); info["rerun_max_abs_diff"] = float((w - w2).abs().max())
else:
    ref = torch.load("gpurun_out/hv_ref.pt").to(dev)
    info["rel_err_vs_stock_fp32"] = float((w - ref).abs().max() / ref.abs().max())
    info["rel_l2_err"] = float((w - ref).norm() / ref.norm())
print(json.dumps(info))
json.dump(info, open(f"gpurun_out/hvp_emul_{mode}_{os.environ.get('HLV_TAG','x')}.json", "w"))
