"""How far apart are valid evaluations of the SAME Lanczos run at the BASELINE size?  (evidence tool)

GPT-2 124M, m=100, full reorthogonalisation, global batch 8x512.  The fp32 double-backward is deterministic but
not linear at the last bit: two runs whose Lanczos vectors differ by one ulp see operator outputs ~5e-6 apart, and
the recurrence amplifies that where the extreme Ritz values converge.  This script measures that sensitivity as an
ENSEMBLE instead of a single pair:

  ours_fused / ours_unfused ........ the CUDA path, CGS2 as 3 passes (fused middle pass) or 4 passes
  oracle_f64 / oracle_f32 .......... oracle.lanczos_cgs2 over the reference's HVP formulation, recurrence in fp64 / fp32 (CUDA ops)
  *_p .............................. the same four from v0' = v0 moved by ONE ulp in every element

and prints every pairwise max |d alpha|, |d beta| over iterations, normalised by max|T|, plus the per-iteration
size of the second Gram-Schmidt correction.  Dev tool: imports oracle/ as the checker (never the product path).

  python scripts/parity_ensemble.py [--iters 100] > gpurun_out/parity_ensemble.json"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hessian_llm_vision_b200 as hlv  # noqa: E402
import oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--global-batch", type=int, default=8)
    ap.add_argument("--no-perturbed", action="store_true")
    args = ap.parse_args()
    from transformers import GPT2Config, GPT2LMHeadModel
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config(vocab_size=50257, n_positions=512, attn_implementation="eager")).eval().to(dev)
    n = sum(p.numel() for p in model.parameters())
    m = args.iters
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, 50257, (args.global_batch, 512), generator=g).to(dev)
    torch.manual_seed(7)
    v0 = torch.randn(n)
    v0 /= v0.double().norm().float()
    v0 = v0.to(dev)
    v0p = torch.nextafter(v0, torch.full_like(v0, float("inf")))        # one ulp up, every element
    runs, secs, diag = {}, {}, {}

    def ours(tag, v, fused):
        op = hlv.HessianVectorProduct(model, [ids])
        c2 = []

        def hook(j, eng):
            if j >= 4:
                cur = eng.coef2 if eng.cgs_passes == 2 else eng.coef
                c2.append(float(cur[: j + 1].abs().max() / eng.norm2.sqrt()))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = hlv.lanczos(op, m, v, reorth="full", fused_cgs=fused, on_iteration=hook)
        torch.cuda.synchronize(); secs[tag] = time.perf_counter() - t0
        runs[tag] = (res.alphas.double().cpu(), res.betas.double().cpu())
        if c2:
            diag[tag + "_max_c2_over_norm_w"] = c2
        del res, op
        torch.cuda.empty_cache()

    def orc(tag, v, dtype):
        def ref_hvp(x):
            return oracle.hess_vec_dataset(x.float().to(dev), [ids], model, weights=[1.0]).to(dtype)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ref = oracle.lanczos_cgs2(ref_hvp, v.to(dtype), m, reorth="full", dtype=dtype)
        torch.cuda.synchronize(); secs[tag] = time.perf_counter() - t0
        runs[tag] = (ref["alphas"].double().cpu(), ref["betas"].double().cpu())
        scale = float(ref["T"].abs().max())
        del ref
        torch.cuda.empty_cache()
        return scale

    ours("ours_fused", v0, True)
    ours("ours_unfused", v0, False)
    scale = orc("oracle_f64", v0, torch.float64)
    orc("oracle_f32", v0, torch.float32)
    if not args.no_perturbed:
        ours("ours_fused_p", v0p, True)
        ours("ours_unfused_p", v0p, False)
        orc("oracle_f64_p", v0p, torch.float64)
        orc("oracle_f32_p", v0p, torch.float32)
    names = list(runs)
    pair = {}
    for i, x in enumerate(names):
        for y in names[i + 1:]:
            da = (runs[x][0] - runs[y][0]).abs() / scale
            db = (runs[x][1] - runs[y][1]).abs() / scale
            pair[f"{x} vs {y}"] = {"alpha": float(da.max()), "beta": float(db.max()), "alpha_argmax": int(da.argmax()),
                                   "alpha_by_decile": [float(da[k: k + max(m // 10, 1)].max()) for k in range(0, m, max(m // 10, 1))]}
    out = {"what": __doc__.split("\n")[0], "P": n, "iters": m, "T_abs_max": scale, "seconds": secs, "pairwise": pair, "diagnostics": diag,
           "alphas": {k: v[0].tolist() for k, v in runs.items()}, "betas": {k: v[1].tolist() for k, v in runs.items()}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
