"""Full-size parity run (BASELINE config 2): GPT-2 124M, m=100, full reorthogonalisation -- this repo's CUDA path
against the oracle recurrence (oracle.lanczos_cgs2, the restatement of lanczostrain_hand.py:171-203 + full
reorthogonalisation) driven by the reference's own HVP formulation (`sum(v*g).backward(); torch.cat`,
gpt2_hessian_cpu.py:75-109), same model, same token batch, same v0.  Three arithmetic settings of the oracle:
  f64      recurrence in float64 (the exact-arithmetic reading of the algorithm; HVP still fp32 on fp32-rounded v)
  f32_gpu  recurrence in float32 with torch CUDA ops -- how the reference's hand loop / gpytorch(device='cuda') runs
  f32_cpu  (--cpu) recurrence in float32 with torch CPU ops -- the gpt2_hessian_cpu.py shape (`.cpu()` every iteration)
North-star bars: alpha/beta within 1e-5 relative per iteration, top-k Ritz values within 1e-4 relative.
Dev/evidence tool: imports oracle/ as the checker (never the product path).

  python scripts/full_size_parity.py [--iters 100] [--cpu] > gpurun_out/full_size_parity.json"""
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
    ap.add_argument("--operator", default="gpt2", choices=["gpt2", "diag"],
                    help="diag: H = diag(d) at the same n -- an operator WITHOUT the fp32 double-backward's rounding noise, to separate "
                         "the recurrence kernels' own error from the HVP's")
    ap.add_argument("--reorth-tol", type=float, default=None, help="also run the CUDA path with the conditional last pass (reorth_tol)")
    ap.add_argument("--skip-f32-gpu", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="also run the fp32 torch-CPU recurrence (needs (m+6)*4n bytes of host RAM, minutes)")
    args = ap.parse_args()
    import psutil
    from transformers import GPT2Config, GPT2LMHeadModel
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config(vocab_size=50257, n_positions=512, attn_implementation="eager")).eval().to(dev)
    n = sum(p.numel() for p in model.parameters())
    m = args.iters
    avail = psutil.virtual_memory().available
    if args.cpu and avail < (m + 6) * n * 4 * 1.15:
        m = max(10, int(avail / 1.15 / (n * 4)) - 6)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, 50257, (args.global_batch, 512), generator=g).to(dev)
    torch.manual_seed(7)
    v0 = torch.randn(n)
    v0 /= v0.double().norm().float()       # float64 reduction: torch's CPU float32 norm is 1.4% off at this length

    if args.operator == "diag":
        gd = torch.Generator(device=dev).manual_seed(3)
        diag = torch.randn(n, device=dev, generator=gd) * 2
        op = lambda v: diag * v.reshape(-1)
    else:
        op = hlv.HessianVectorProduct(model, [ids])
    # ---- this repo: everything on the device through libhlv ----
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = hlv.lanczos(op, m, v0.to(dev), reorth="full")
    torch.cuda.synchronize(); t_ours = time.perf_counter() - t0
    a, b = res.alphas.double(), res.betas.double()
    ev = res.eigvals.double()
    del res
    torch.cuda.empty_cache()
    # a second run of THIS path with a different (equally valid) reduction order: CGS2 as 4 separate passes
    res_u = hlv.lanczos(op, m, v0.to(dev), reorth="full", fused_cgs=False)
    a_u, b_u = res_u.alphas.double(), res_u.betas.double()
    del res_u, op
    torch.cuda.empty_cache()
    coeffs = {"ours": (a, b), "ours_unfused": (a_u, b_u)}
    if args.reorth_tol is not None:
        op_t = (lambda v: diag * v.reshape(-1)) if args.operator == "diag" else hlv.HessianVectorProduct(model, [ids])
        res_t = hlv.lanczos(op_t, m, v0.to(dev), reorth="full", reorth_tol=args.reorth_tol)
        coeffs[f"ours_reorth_tol_{args.reorth_tol:g}"] = (res_t.alphas.double(), res_t.betas.double())
        ev_t = res_t.eigvals.double()
        Qt = res_t.Q
        G = torch.zeros(m, m, dtype=torch.float64, device=dev)
        for c0 in range(0, Qt.shape[1], 1 << 22):
            Qc = Qt[:, c0: c0 + (1 << 22)].double()
            G += Qc @ Qc.t()
        orth_t = float((G - torch.eye(m, dtype=torch.float64, device=dev)).abs().max())
        del res_t, op_t, Qt, G
        torch.cuda.empty_cache()

    def compare(ref, seconds):
        a_ref, b_ref = ref["alphas"].double().cpu(), ref["betas"].double().cpu()
        ev_ref = torch.linalg.eigvalsh(ref["T"].double().cpu())
        scale = float(ref["T"].abs().max())
        err_a, err_b = (a - a_ref).abs() / scale, (b - b_ref).abs() / scale
        k = min(10, m)
        top_rel = ((ev[-k:] - ev_ref[-k:]).abs() / ev_ref[-k:].abs()).tolist()
        step = max(m // 10, 1)
        return {"T_abs_max": scale, "alpha_max_rel_err": float(err_a.max()), "beta_max_rel_err": float(err_b.max()),
                "alpha_rel_err_by_decile": [float(err_a[i: i + step].max()) for i in range(0, m, step)],
                "beta_rel_err_by_decile": [float(err_b[i: i + step].max()) for i in range(0, m, step)],
                "ritz_top10_oracle": ev_ref[-k:].tolist(), "ritz_top10_rel_err": top_rel,
                "ritz_all_max_abs_err_over_scale": float((ev - ev_ref).abs().max() / scale),
                "pass": bool(float(err_a.max()) < 1e-5 and float(err_b.max()) < 1e-5 and max(top_rel) < 1e-4),
                "seconds": seconds, "iterations_per_s": m / seconds}

    def ref_hvp(v):                                   # the reference's formulation, fp32, on the GPU
        if args.operator == "diag":
            return diag * v.float().to(dev)
        return oracle.hess_vec_dataset(v.float().to(dev), [ids], model, weights=[1.0])

    out = {"what": "GPT-2 124M, full CGS2 reorthogonalisation: CUDA path vs the oracle recurrence over the reference's HVP, same v0 and batch",
           "operator": args.operator, "P": n, "iters": m, "iters_requested": args.iters, "global_batch": args.global_batch,
           "bars": {"alpha_beta_per_iteration": 1e-5, "ritz_top_k": 1e-4},
           "ritz_top10_ours": ev[-min(10, m):].tolist(), "seconds_cuda_path": t_ours, "iterations_per_s_cuda_path": m / t_ours}
    settings = (("f64", torch.float64, dev),) + (() if args.skip_f32_gpu else (("f32_gpu", torch.float32, dev),))
    for name, dtype, where in settings + ((("f32_cpu", torch.float32, "cpu"),) if args.cpu else ()):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ref = oracle.lanczos_cgs2(lambda v: ref_hvp(v).to(device=where, dtype=dtype), v0.to(device=where, dtype=dtype), m,
                                  reorth="full", dtype=dtype)
        torch.cuda.synchronize()
        out["oracle_" + name] = compare(ref, time.perf_counter() - t0)
        coeffs["oracle_" + name] = (ref["alphas"].double().cpu(), ref["betas"].double().cpu())
        if name == "f64" and args.reorth_tol is not None:
            ev64 = torch.linalg.eigvalsh(ref["T"].double().cpu())
            out["ours_reorth_tol"] = {"tol": args.reorth_tol, "max_abs_QQt_minus_I": orth_t,
                                      "ritz_top10_rel_err_vs_f64": ((ev_t[-10:] - ev64[-10:]).abs() / ev64[-10:].abs()).tolist(),
                                      "ritz_all_max_abs_err_over_scale": float((ev_t - ev64).abs().max() / float(ref["T"].abs().max()))}
        del ref
        torch.cuda.empty_cache()
    # how far apart are two valid evaluations of the SAME algorithm?  (the sensitivity floor of alpha/beta at this size)
    scale = out["oracle_f64"]["T_abs_max"]
    names = list(coeffs)
    out["pairwise_max_rel_diff_alpha_beta"] = {
        f"{x} vs {y}": [float((coeffs[x][0] - coeffs[y][0]).abs().max() / scale), float((coeffs[x][1] - coeffs[y][1]).abs().max() / scale)]
        for i, x in enumerate(names) for y in names[i + 1:]}
    out["pass"] = out["oracle_f64"]["pass"] and out.get("oracle_f32_gpu", {"pass": True})["pass"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
