"""GPU probe (dev tool): accuracy and speed of the 3xTF32 GEMM override (fp32_tc) for the GPT-2 HVP.
Truth = the same HVP in float64; compares stock fp32 (SIMT sgemm), 3xTF32 and plain TF32 against it."""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import hessian_llm_vision_b200 as hlv
from hessian_llm_vision_b200 import fp32_tc

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model, cfg = bench.build_model(False)
model.to(dev)
n = sum(p.numel() for p in model.parameters())
ids = bench.make_tokens(cfg, B, B, 512)[0].to(dev)
g = torch.Generator(device=dev).manual_seed(3)
v = torch.randn(n, device=dev, generator=g)
v /= v.norm()
out = {"B": B}


def run(tag, reps=4):
    op = hlv.HessianVectorProduct(model, [ids])
    w = torch.empty(n, device=dev)
    for _ in range(2):
        op.accumulate_into(v, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        op.accumulate_into(v, w)
    e1.record()
    torch.cuda.synchronize()
    out[tag] = {"gpu_ms": e0.elapsed_time(e1) / reps, "wall_ms": (time.perf_counter() - t0) * 1e3 / reps}
    return w.double()


# float64 truth (pieces concatenated with torch.cat: the libhlv gather is fp32-only)
m64 = copy.deepcopy(model).double()
params = list(m64.parameters())
loss = m64(input_ids=ids, labels=ids).loss
grads = torch.autograd.grad(loss, params, create_graph=True)
views = [s.view_as(p) for s, p in zip(torch.split(v.double(), [p.numel() for p in params]), params)]
hv64 = torch.cat([h.reshape(-1) for h in torch.autograd.grad(grads, params, grad_outputs=views)])
del m64, grads, loss
torch.cuda.empty_cache()


def err(w):
    return {"rel_l2": float((w - hv64).norm() / hv64.norm()), "rel_max": float((w - hv64).abs().max() / hv64.abs().max())}


w_stock = run("stock_fp32")
out["stock_fp32"].update(err(w_stock))
fp32_tc.enable()
w_3x = run("3xtf32")
out["3xtf32"].update(err(w_3x))
out["3xtf32"]["gemm_calls_overridden"] = fp32_tc.calls
fp32_tc.disable()
torch.backends.cuda.matmul.allow_tf32 = True
w_tf = run("tf32")
out["tf32"].update(err(w_tf))
torch.backends.cuda.matmul.allow_tf32 = False
w_again = run("stock_after_disable")
out["stock_after_disable"]["bitwise_equal_to_stock"] = bool(torch.equal(w_again, w_stock))
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/hvp_3xtf32_B{B}.json", "w"), indent=1)
