"""Small, fast exercise of every libhlv kernel on ragged sizes -- the target for
`compute-sanitizer --tool memcheck` (one tool per gpurun call, smallest case that shows what is needed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hessian_llm_vision_b200 import kernels as K

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ws = K.Workspace(dev, max_rows=130)
f64 = dict(dtype=torch.float64, device=dev)
for n in (1, 7, 2049, 70_003):
    sizes = [n // 3, 0, n - n // 3] if n > 2 else [n]
    ts = [torch.randn(s, device=dev, generator=g) for s in sizes]
    v = torch.randn(n, device=dev, generator=g)
    dst = torch.empty(n, device=dev)
    dot = torch.zeros(1, **f64)
    K.gather(ts, dst, dot_with=v, dot_out=dot, ws=ws)
    K.gather(ts, dst, scale=0.5, accumulate=True)
    K.scatter(dst, [torch.empty_like(t) for t in ts])
    K.dot(dst, v, dot, ws)
    nrm = torch.zeros(1, **f64)
    K.lanczos_update(dst, v, v.clone(), dot, dot.clone(), nrm, ws)
    K.lanczos_update(dst, v, None, dot, None, nrm, ws)
    n8 = (n + 7) // 8 * 8
    vo = torch.zeros(n8, device=dev)[:n]
    r16 = torch.zeros(n8, dtype=torch.bfloat16, device=dev)[:n]
    K.normalize_store(dst, nrm, dot, vo, r16, 0.0, torch.full((1,), -1, dtype=torch.int32, device=dev), 0)
    for dt in (torch.float32, torch.bfloat16):
        for rows in (1, 9, 100):
            V = torch.zeros(rows, n8, dtype=dt, device=dev)
            V[:, :n] = (torch.randn(rows, n, device=dev, generator=g) / max(n, 1) ** 0.5).to(dt)
            c = torch.zeros(rows, **f64)
            c2 = torch.zeros(rows, **f64)
            w = torch.randn(n, device=dev, generator=g)
            K.cgs_project(V, rows, w, c, ws)
            K.cgs_update(V, rows, c, w, nrm, ws)
            K.cgs_update_project(V, rows, c, w, c2, nrm, ws)
            Y = torch.randn(rows, 3, device=dev, generator=g)
            out = torch.empty(3, n8, device=dev)
            K.ritz_vectors(V, rows, Y, out, n)
    Vf = torch.randn(5, n8, device=dev, generator=g)
    K.vector_adjust(v, Vf[:, :n] if n8 != n else Vf, torch.rand(5, device=dev, generator=g) + 0.5, v.clone(), 1e-2, ws)
torch.cuda.synchronize()
print("sanitize target ok")
