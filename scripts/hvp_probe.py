"""GPU probe (dev tool): time the GPT-2 124M HVP per micro-batch, memory, cached-graph variant."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import hessian_llm_vision_b200 as hlv

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
model, cfg = bench.build_model(False)
model.to(dev)
n = sum(p.numel() for p in model.parameters())
out = {"P": n}
v = torch.randn(n, device=dev); v /= v.norm()
w = torch.empty(n, device=dev)
for B in (1, 2, 4, 8, 16):
    ids = bench.make_tokens(cfg, B, B, 512)[0].to(dev)
    for cache in (False, True):
        torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
        op = hlv.HessianVectorProduct(model, [ids], cache_graph=cache)
        for _ in range(2):
            op.accumulate_into(v, w)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            op.accumulate_into(v, w)
        e1.record(); torch.cuda.synchronize()
        out[f"B{B}_cache{int(cache)}"] = {"gpu_ms": e0.elapsed_time(e1) / 5, "wall_ms": (time.perf_counter() - t0) * 200,
                                          "peak_gb": torch.cuda.max_memory_allocated() / 2**30}
        op.clear_cache(); del op
        print(B, cache, out[f"B{B}_cache{int(cache)}"], flush=True)
json.dump(out, open("gpurun_out/hvp_probe.json", "w"), indent=1)
