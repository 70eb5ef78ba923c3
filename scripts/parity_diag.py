"""Diagnostic: where do the CUDA path and the reference-shape oracle part ways at GPT-2 size?"""
import copy, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hessian_llm_vision_b200 as hlv
import oracle
from transformers import GPT2Config, GPT2LMHeadModel

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
model = GPT2LMHeadModel(GPT2Config(vocab_size=50257, n_positions=512, attn_implementation="eager")).eval().to(dev)
n = sum(p.numel() for p in model.parameters())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = torch.Generator().manual_seed(1234)
ids = torch.randint(0, 50257, (8, 512), generator=g)[:B].contiguous().to(dev)
torch.manual_seed(7)
v0 = torch.randn(n); v0 /= v0.double().norm().float()   # float64 reduction (the CPU float32 norm is 1.4% off here)
v0d = v0.to(dev)
out = {"B": B, "m": m}

# --- one HVP three ways + fp64 ground truth
op = hlv.HessianVectorProduct(model, [ids])
hv_a = op(v0d).double()
hv_a2 = op(v0d).double()
hv_b = oracle.hess_vec_dataset(v0d, [ids], model, weights=[1.0]).double()
model64 = copy.deepcopy(model).double()
hv_t = oracle.hess_vec_dataset(v0d.double(), [ids], model64, weights=[1.0]).double()
del model64
torch.cuda.empty_cache()
nt = float(hv_t.norm())
out["hvp"] = {"norm_fp64": nt, "ours_vs_fp64": float((hv_a - hv_t).norm()) / nt, "oracle_vs_fp64": float((hv_b - hv_t).norm()) / nt,
              "ours_vs_oracle": float((hv_a - hv_b).norm()) / nt, "ours_run_to_run": float((hv_a - hv_a2).norm()) / nt,
              "alpha0_fp64": float(hv_t @ v0d.double()), "alpha0_ours": float(hv_a @ v0d.double()), "alpha0_oracle": float(hv_b @ v0d.double())}
del hv_a, hv_a2, hv_b, hv_t

# --- short Lanczos runs
res = hlv.lanczos(op, m, v0d, reorth="full")
a, b = res.alphas.double(), res.betas.double()
del res
torch.cuda.empty_cache()
mv32 = lambda v: oracle.hess_vec_dataset(v.to(dev), [ids], model, weights=[1.0]).cpu()
r32 = oracle.lanczos_cgs2(mv32, v0, m, reorth="full")
mv64 = lambda v: oracle.hess_vec_dataset(v.float().to(dev), [ids], model, weights=[1.0]).double().cpu()
r64 = oracle.lanczos_cgs2(mv64, v0.double(), m, reorth="full", dtype=torch.float64)     # fp64 recurrence, fp32 HVP on fp32-rounded v
out["alphas"] = {"ours": a.tolist(), "oracle_f32": r32["alphas"].double().tolist(), "oracle_f64rec": r64["alphas"].tolist()}
out["betas"] = {"ours": b.tolist(), "oracle_f32": r32["betas"].double().tolist(), "oracle_f64rec": r64["betas"].tolist()}
print(json.dumps(out))
