"""GPU micro-benchmark (dev tool): the three Gram-Schmidt kernels alone at GPT-2 size.
  python scripts/cgs_bench.py [f32|bf16] [rows,rows,...]
Prints achieved GB/s (algorithmic bytes / CUDA-event time) per kernel and depth."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hessian_llm_vision_b200 import kernels as K

dev = torch.device("cuda:0")
dt = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else torch.bfloat16
rows_list = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "10,26,50,51,100").split(",")]
n = 124_046_592
s = 4 if dt == torch.float32 else 2
R = max(rows_list)
V = torch.empty(R, n, dtype=dt, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for r in range(R):
    V[r].copy_((torch.randn(n, device=dev, generator=g) * n ** -0.5).to(dt))
w0 = torch.randn(n, device=dev, generator=g)
ws = K.Workspace(dev, max_rows=R + 1)
c = torch.zeros(R, dtype=torch.float64, device=dev)
c2 = torch.zeros(R, dtype=torch.float64, device=dev)
nrm = torch.zeros(1, dtype=torch.float64, device=dev)
out = {}


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for rows in rows_list:
    w = w0.clone()
    K.cgs_project(V, rows, w, c, ws)
    t_p = timeit(lambda: K.cgs_project(V, rows, w, c, ws))
    t_u = timeit(lambda: K.cgs_update(V, rows, c, w, nrm, ws))
    res = {"project_gbs": (rows * s + 4) * n / t_p / 1e6, "update_gbs": (rows * s + 8) * n / t_u / 1e6,
           "project_ms": t_p, "update_ms": t_u}
    if rows <= K.fused_max_rows(dt):
        t_f = timeit(lambda: K.cgs_update_project(V, rows, c, w, c2, nrm, ws))
        res.update({"fused_gbs": (rows * s + 8) * n / t_f / 1e6, "fused_ms": t_f,
                    "fused_vs_pair": (t_p + t_u) / t_f})
    out[rows] = res
    print(dt, rows, {k: round(v, 2) for k, v in res.items()}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/cgs_bench_{'f32' if s == 4 else 'bf16'}.json", "w"), indent=1)
