"""Dev probe for the tcgen05 Ritz kernel: structured inputs whose outputs reveal layout / descriptor mistakes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from hessian_llm_vision_b200 import kernels as K
np.set_printoptions(linewidth=250, precision=0, suppress=True)
dev = torch.device("cuda:0")


def run(m, nvec, n, Q, Y, tag):
    out = torch.full((nvec, n), float("nan"), device=dev)
    K.ritz_vectors(Q, m, Y.contiguous(), out, n)
    torch.cuda.synchronize()
    ref = (Y.double().t() @ Q.double()).float()
    print(f"== {tag}: m={m} nvec={nvec} n={n} max err {(out - ref).abs().max().item():.3g} (scale {ref.abs().max().item():.3g})")
    return out.cpu().numpy(), ref.cpu().numpy()


for m in (8, 16, 24):
    nvec, n = 16, 128
    Q = (torch.arange(m, device=dev).float()[:, None] * 1000 + torch.arange(n, device=dev).float()[None, :]).contiguous()
    Y = torch.zeros(m, nvec, device=dev)
    k = min(m, nvec)
    Y[torch.arange(k), torch.arange(k)] = 1.0
    out, ref = run(m, nvec, n, Q, Y, "identity Y, Q[i,x]=1000i+x")
    print("out[:, 0:12]"); print(out[:, 0:12])
    print("out[0:4, 8:40]"); print(out[0:4, 8:40])
m, nvec, n = 8, 16, 128
Q = (torch.arange(m, device=dev).float()[:, None] * 1000 + torch.arange(n, device=dev).float()[None, :]).contiguous()
for i0, r0 in ((0, 0), (1, 0), (4, 0), (0, 1), (0, 9), (5, 9)):
    Y = torch.zeros(m, nvec, device=dev)
    Y[i0, r0] = 1.0
    out, ref = run(m, nvec, n, Q, Y, f"Y[{i0},{r0}]=1")
    rows = sorted(set(np.argwhere(out != 0)[:, 0].tolist()))
    print("   nonzero output rows:", rows, "; first:", out[rows[0], :10] if rows else None)
