"""GPU probe (dev tool): time the GPT-2 124M HVP per micro-batch: eager launches, cached first-backward
graph, and CUDA-graph replay; checks the variants agree."""
import json
import os
import sys
import time

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
v = torch.randn(n, device=dev)
v /= v.norm()
w = torch.empty(n, device=dev)
batches = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,2,4,8".split(","))]


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps


for B in batches:
    ids = bench.make_tokens(cfg, B, B, 512)[0].to(dev)
    ref = None
    for variant in ("eager", "cache", "cudagraph"):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        try:
            # cudagraph: first-backward graph built eagerly once (HF forward code is not capturable: it copies
            # CPU scalars to the device), the second backward + gather are captured and replayed
            op = hlv.HessianVectorProduct(model, [ids], cache_graph=(variant in ("cache", "cudagraph")))
            run_op = op.capture() if variant == "cudagraph" else op
            gpu_ms, wall_ms = timeit(lambda: run_op.accumulate_into(v, w))
            if ref is None:
                ref = w.clone()
                err = 0.0
            else:
                err = float((w - ref).abs().max() / ref.abs().max())
            out[f"B{B}_{variant}"] = {"gpu_ms": gpu_ms, "wall_ms": wall_ms, "peak_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                                      "rel_err_vs_eager": err}
        except Exception as e:  # noqa: BLE001
            import traceback
            out[f"B{B}_{variant}"] = {"error": repr(e)[:400], "traceback": traceback.format_exc()[-3000:]}
        print(B, variant, out[f"B{B}_{variant}"], flush=True)
        try:
            op.clear_cache()
            del run_op, op
        except Exception:  # noqa: BLE001
            pass
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/hvp_probe.json", "w"), indent=1)
