"""GPU probe (dev tool): torch.profiler breakdown of one GPT-2 124M HVP (B=8) by CUDA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import hessian_llm_vision_b200 as hlv

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
model, cfg = bench.build_model(False)
model.to(dev)
n = sum(p.numel() for p in model.parameters())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ids = bench.make_tokens(cfg, B, B, 512)[0].to(dev)
v = torch.randn(n, device=dev); v /= v.norm()
w = torch.empty(n, device=dev)
op = hlv.HessianVectorProduct(model, [ids])
for _ in range(2):
    op.accumulate_into(v, w)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    op.accumulate_into(v, w)
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90)
open(f"gpurun_out/hvp_profile_B{B}.txt", "w").write(tab)
print(tab[-6000:])
