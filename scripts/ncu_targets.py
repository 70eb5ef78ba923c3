"""ncu target (dev tool): every libhlv kernel of the recurrence once or twice at GPT-2 size, nothing else.
  python scripts/ncu_targets.py [rows]
  ncu --set full --clock-control none --import-source on -k regex:'cgs|ritz|normalize|lanczos|multi_tensor|reduce_scatter' -o out python scripts/ncu_targets.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hessian_llm_vision_b200 import kernels as K

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = 124_046_592
g = torch.Generator(device=dev).manual_seed(0)
ws = K.Workspace(dev, max_rows=rows + 2)
c = torch.zeros(rows, dtype=torch.float64, device=dev)
c2 = torch.zeros(rows, dtype=torch.float64, device=dev)
nrm = torch.zeros(1, dtype=torch.float64, device=dev)
beta = torch.ones(1, dtype=torch.float64, device=dev)
alpha = torch.full((1,), 0.5, dtype=torch.float64, device=dev)
w = torch.randn(n, device=dev, generator=g)
for dt in (torch.float32, torch.bfloat16):
    V = torch.empty(rows, n, dtype=dt, device=dev)
    for r in range(rows):
        V[r].copy_((torch.randn(n, device=dev, generator=g) * n ** -0.5).to(dt))
    vj = V[rows - 1].float().contiguous() if dt != torch.float32 else V[rows - 1]
    vo = V[rows - 2].float().contiguous() if dt != torch.float32 else V[rows - 2]
    for rep in range(2):
        K.x_update_project(None, V, rows, w, vj, vo, alpha, beta, c, ws)        # three-term update + first projection
        K.cgs_update_project(V, rows, c, w, c2, nrm, ws)                         # fused middle pass (TMA slab)
        K.cgs_update(V, rows, c2, w, nrm, ws)                                    # last update + |w|^2
    if dt == torch.float32:
        K.cgs_project(V, rows, w, c, ws)
        out = torch.empty(n, device=dev)
        K.normalize_store(w, nrm, beta, out)
        K.lanczos_update(w, vj, vo, alpha, beta, nrm, ws)
        K.dot(w, vj, alpha, ws)
        del out
        Y = torch.linalg.qr(torch.randn(rows, rows, device=dev, generator=g))[0].contiguous()
        try:
            outv = torch.empty(rows, n, device=dev)
            K.ritz_vectors(V, rows, Y, outv, n)                                   # tcgen05 pass
            K.ritz_vectors(V, rows, Y, outv, n)
            del outv
        except torch.OutOfMemoryError:
            pass
    del V
    torch.cuda.empty_cache()
torch.cuda.synchronize()
print("ok")
