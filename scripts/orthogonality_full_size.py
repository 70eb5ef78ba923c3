"""Dev tool: max|QQ^T - I| of the stored basis after the GPT-2 124M m=100 full-reorthogonalisation run (default path)."""
import os, sys, json, torch
sys.path.insert(0, "/root/repo")
import hessian_llm_vision_b200 as hlv
from transformers import GPT2Config, GPT2LMHeadModel
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
model = GPT2LMHeadModel(GPT2Config(vocab_size=50257, n_positions=512, attn_implementation="eager")).eval().to(dev)
n = sum(p.numel() for p in model.parameters())
ids = torch.randint(0, 50257, (8, 512), generator=torch.Generator().manual_seed(1234)).to(dev)
torch.manual_seed(7); v0 = torch.randn(n); v0 /= v0.double().norm().float()   # NOT v0.norm(): CPU float32 norm is 1.4% off at this length
res = hlv.lanczos(hlv.HessianVectorProduct(model, [ids]).capture(), 100, v0.to(dev), reorth="full")
Q = res.Q; m = res.m
G = torch.zeros(m, m, dtype=torch.float64, device=dev)
for c0 in range(0, Q.shape[1], 1 << 22):
    Qc = Q[:, c0: c0 + (1 << 22)].double(); G += Qc @ Qc.t()
print(json.dumps({"m": m, "max_abs_QQt_minus_I": float((G - torch.eye(m, dtype=torch.float64, device=dev)).abs().max()), "ritz_top3": res.eigvals[-3:].tolist()}))
