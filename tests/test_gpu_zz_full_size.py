"""-m gpu, BASELINE configuration 2 at FULL size: GPT-2 124M (P = 124,046,592), m = 100, full two-pass classical
Gram-Schmidt, global batch 8 x 512, same v0 -- the CUDA path through its default settings against the oracle recurrence
(oracle.lanczos_cgs2: lanczostrain_hand.py:171-203 order + Discrepancy.ipynb cell 1:36-54 reorthogonalisation) driven by
the reference's own HVP formulation (gpt2_hessian_cpu.py:75-109).  Runs last (file name) and takes ~2 minutes on a B200.

Definition of "relative" used by every parity test in this repository: |x_ours - x_oracle| / max|T|, per iteration
(alpha and beta of a Lanczos run span orders of magnitude and alpha crosses zero, so a per-entry quotient is not
meaningful; max|T| is the operator-norm scale the north star's 1e-5 refers to).

What the bar is, and why (profiles/r02_parity_ensemble_gpt2_m100.json): the fp32 double-backward is deterministic but not
linear at the last bit, so two valid evaluations of the SAME algorithm whose Lanczos vectors differ by one ulp drift
apart: the reference's own fp32 loop moves by 2.0e-5 when v0 moves by ONE ulp, the float64 recurrence by 4.6e-6, and
fp32-vs-float64 oracle runs differ by 1.2e-5.  The test therefore measures that floor in the same run -- the distance
between the oracle evaluated in float64 and in float32 (torch CUDA ops, how the reference's hand loop runs) -- and asks
the CUDA path to be within 1e-5 + floor of the float64 recurrence, the same rule as test_full_reorth_vs_oracle.
"""
import json
import os

import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

P_GPT2 = 124_046_592


@pytest.fixture(scope="module")
def gpt2(cuda_dev, libhlv):
    from transformers import GPT2Config, GPT2LMHeadModel
    free, _ = torch.cuda.mem_get_info(cuda_dev)
    if free < 150e9:
        pytest.skip("needs a 180 GB device")
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config(vocab_size=50257, n_positions=512, attn_implementation="eager")).eval().to(cuda_dev)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, 50257, (8, 512), generator=g).to(cuda_dev)
    torch.manual_seed(7)
    v0 = torch.randn(P_GPT2)
    v0 /= v0.double().norm().float()           # float64 reduction: torch's CPU float32 norm is 1.4% off at this length
    return model, ids, v0.to(cuda_dev)


def _record(name, payload):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, name), "w") as f:
            json.dump(payload, f, indent=1)


def test_kept_first_backward_graph_is_bit_identical_full_size(gpt2, cuda_dev):
    """The library default keeps the v-independent half of the double-backward (forward + first backward,
    gpt2_hessian_cpu.py:94-102) for the whole run; an application then performs only the second backward.  Same kernels
    on the same saved tensors: H v must be BIT-identical to rebuilding everything, eagerly and replayed from CUDA graphs."""
    import hessian_llm_vision_b200 as hlv
    model, ids, v0 = gpt2
    assert sum(p.numel() for p in model.parameters()) == P_GPT2
    rebuild = hlv.HessianVectorProduct(model, [ids], cache_graph=False)
    keep = hlv.HessianVectorProduct(model, [ids])                  # default: "auto"
    a = rebuild(v0)
    b1 = keep(v0)
    assert keep.cache_graph is True and keep.first_backward_builds == 1 and keep.cached_bytes > 0
    v1 = torch.roll(v0, 12345)
    b2 = keep(v1)
    assert keep.first_backward_builds == 1                         # second application: no forward, no first backward
    assert torch.equal(a, b1)
    assert torch.equal(rebuild(v1), b2)
    ref = oracle.hess_vec_dataset(v0, [ids], model, weights=[1.0])  # the reference's formulation, sum(v*g).backward()
    assert torch.equal(ref, a)
    g = keep.capture()                                             # two graphs: first half once, second half per application
    c1, c2, c3 = g(v0), g(v1), g(v0)
    assert g.first_replays == 1
    assert torch.equal(c1, a) and torch.equal(c2, b2) and torch.equal(c3, a)
    g.invalidate()
    assert torch.equal(g(v1), b2) and g.first_replays == 2
    _record("r02_hvp_keep_first_backward.json", {"bit_identical": True, "kept_graph_bytes": keep.cached_bytes})


def test_config2_full_size_parity(gpt2, cuda_dev):
    import hessian_llm_vision_b200 as hlv
    model, ids, v0 = gpt2
    m = 100
    op = hlv.HessianVectorProduct(model, [ids])
    res = hlv.lanczos(op, m, v0, reorth="full")                    # every default: fused TMA pass, kept first-backward graph
    a, b, ev = res.alphas.double().cpu(), res.betas.double().cpu(), res.eigvals.double().cpu()
    Q = res.Q
    G = torch.zeros(m, m, dtype=torch.float64, device=cuda_dev)
    for c0 in range(0, Q.shape[1], 1 << 22):
        Qc = Q[:, c0: c0 + (1 << 22)].double()
        G += Qc @ Qc.t()
    orth = float((G - torch.eye(m, dtype=torch.float64, device=cuda_dev)).abs().max())
    del res, op, Q, G
    torch.cuda.empty_cache()

    def ref_hvp(dtype):
        return lambda v: oracle.hess_vec_dataset(v.float(), [ids], model, weights=[1.0]).to(dtype)
    runs = {}
    for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        ref = oracle.lanczos_cgs2(ref_hvp(dtype), v0.to(dtype), m, reorth="full", dtype=dtype)
        runs[name] = (ref["alphas"].double().cpu(), ref["betas"].double().cpu(), torch.linalg.eigvalsh(ref["T"].double().cpu()),
                      float(ref["T"].abs().max()))
        del ref
        torch.cuda.empty_cache()
    a64, b64, ev64, scale = runs["f64"]
    a32, b32, ev32, _ = runs["f32"]
    rel = lambda x, y: float((x - y).abs().max()) / scale
    floor_a, floor_b = rel(a64, a32), rel(b64, b32)
    err = {"alpha_vs_f64": rel(a, a64), "beta_vs_f64": rel(b, b64), "alpha_vs_f32": rel(a, a32), "beta_vs_f32": rel(b, b32),
           "floor_alpha_f64_vs_f32": floor_a, "floor_beta_f64_vs_f32": floor_b, "T_abs_max": scale, "max_abs_QQt_minus_I": orth,
           "ritz_top10_rel_vs_f64": float(((ev[-10:] - ev64[-10:]) / ev64[-10:]).abs().max()),
           "ritz_all_abs_over_scale_vs_f64": rel(ev, ev64), "bar": "1e-5 + floor (alpha, beta vs f64); 1e-4 (top-10 Ritz, relative)"}
    _record("r02_full_size_parity_test.json", err)
    print(json.dumps(err))
    assert err["alpha_vs_f64"] < 1e-5 + floor_a, err               # north star: alpha/beta within 1e-5 relative per iteration
    assert err["beta_vs_f64"] < 1e-5 + floor_b, err
    assert err["ritz_top10_rel_vs_f64"] < 1e-4, err                # top-k Ritz values within 1e-4 relative
    assert err["ritz_all_abs_over_scale_vs_f64"] < 1e-5, err
    assert orth < 5e-6, err                                        # the stored rows are orthonormal at fp32 working precision


def _oracle_pair(hvp32, v0, m, reorth):
    """The oracle recurrence in float64 and in float32 (torch CUDA ops) over the same fp32 operator: (f64, f32, scale)."""
    out = {}
    for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        ref = oracle.lanczos_cgs2(lambda v: hvp32(v.float()).to(dtype), v0.to(dtype), m, reorth=reorth, dtype=dtype)
        out[name] = (ref["alphas"].double().cpu(), ref["betas"].double().cpu(), torch.linalg.eigvalsh(ref["T"].double().cpu()),
                     float(ref["T"].abs().max()))
        del ref
        torch.cuda.empty_cache()
    return out["f64"], out["f32"], out["f64"][3]


def test_config1_full_size_dataset_hvp_no_reorth(gpt2, cuda_dev):
    """BASELINE config 1 at FULL size: GPT-2 124M, 25 iterations, NO reorthogonalisation, 20 sequences (= int(1e-4 * 205,328)
    documents) streamed as micro-batches 8 + 8 + 4 with B_i/N weights (gpt2_savehessian.py:143-163, lanczostrain_hand.py:171-203)
    against the oracle over the reference's dataset HVP.  Without reorthogonalisation alpha/beta are rounding-sensitive once the
    first Ritz value has converged, so the bar is per iteration: 1e-5 of max|T| plus the distance between the oracle's own
    float32 and float64 evaluations at that iteration."""
    import hessian_llm_vision_b200 as hlv
    model, _, v0 = gpt2
    g = torch.Generator().manual_seed(4321)
    ids = torch.randint(0, 50257, (20, 512), generator=g).to(cuda_dev)
    batches = [ids[0:8], ids[8:16], ids[16:20]]
    m = 25
    op = hlv.HessianVectorProduct(model, batches)
    assert op.weights == [8 / 20, 8 / 20, 4 / 20]
    res = hlv.lanczos(op, m, v0, reorth=None)
    a, b, ev = res.alphas.double().cpu(), res.betas.double().cpu(), res.eigvals.double().cpu()
    sum_gamma = float(res.gammas.sum())
    op.clear_cache()
    del op, res
    torch.cuda.empty_cache()
    f64, f32, scale = _oracle_pair(lambda v: oracle.hess_vec_dataset(v, batches, model), v0, m, None)
    floor_a, floor_b = (f64[0] - f32[0]).abs() / scale, (f64[1] - f32[1]).abs() / scale
    err_a, err_b = (a - f64[0]).abs() / scale, (b - f64[1]).abs() / scale
    rec = {"alpha_err_max": float(err_a.max()), "beta_err_max": float(err_b.max()), "floor_alpha_max": float(floor_a.max()),
           "floor_beta_max": float(floor_b.max()), "ritz_max_rel": float(abs(ev[-1] - f64[2][-1]) / abs(f64[2][-1])), "T_abs_max": scale}
    _record("r02_config1_full_size_parity_test.json", rec)
    print(json.dumps(rec))
    assert bool((err_a <= 1e-5 + floor_a).all()), rec
    assert bool((err_b <= 1e-5 + floor_b).all()), rec
    assert rec["ritz_max_rel"] < 1e-4, rec                          # the converged extreme Ritz value
    assert abs(sum_gamma - 1.0) < 1e-5


def test_config3_full_size_per_block_spectra(gpt2, cuda_dev):
    """BASELINE config 3 at FULL size for the first and the last transformer block (P_block = 7,087,872 each): one Lanczos run
    per block, operator restricted to the block's parameters (ipynbs/visual-eigen.ipynb cells 10-12), full reorthogonalisation."""
    import hessian_llm_vision_b200 as hlv
    model, ids, _ = gpt2
    m = 20
    blocks = [model.transformer.h[0], model.transformer.h[11]]
    ev, gm = hlv.per_block_spectra(model, [ids], m, blocks=blocks, seed=10)
    out = {}
    for i, blk in enumerate(blocks):
        params = list(blk.parameters())
        nb = sum(p.numel() for p in params)
        assert nb == 7_087_872
        v0 = hlv.probe_vector(nb, 10 + i, cuda_dev)
        f64, f32, scale = _oracle_pair(lambda v: oracle.hess_vec_subset(v, [ids], model, params), v0, m, "full")
        floor = float((f64[2] - f32[2]).abs().max()) / scale
        err = float((ev[i].double() - f64[2]).abs().max()) / scale
        out[f"block{0 if i == 0 else 11}"] = {"ritz_err_over_scale": err, "floor": floor, "ritz_max": float(f64[2][-1]), "sum_gammas": float(gm[i].sum())}
        assert err < 1e-5 + floor, out
        assert abs(float(gm[i].sum()) - 1.0) < 1e-5
    _record("r02_config3_full_size_parity_test.json", out)
    print(json.dumps(out))


def test_conditional_pass_orthogonality_full_size(cuda_dev, libhlv):
    """The gpytorch-style conditional last pass (`reorth_tol=1e-5`) at FULL size (n = 124,046,592, m = 100) on a diagonal
    operator (no HVP noise, seconds to run): the stored rows must stay orthonormal to the rule's own guarantee (no
    projection above tol), T must equal the unconditional two-pass run to working precision, and the run reports how often
    the pass was applied.  (Round 1 had this check only at n = 70,003.)"""
    import hessian_llm_vision_b200 as hlv
    free, _ = torch.cuda.mem_get_info(cuda_dev)
    if free < 120e9:
        pytest.skip("needs ~105 GB of device memory")
    n, m = P_GPT2, 100
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    diag = torch.randn(n, device=cuda_dev, generator=g) * 2
    v0 = hlv.probe_vector(n, 0, cuda_dev)
    op = lambda v: diag * v
    base = hlv.lanczos(op, m, v0, reorth="full")
    a0, b0, ev0 = base.alphas.clone(), base.betas.clone(), base.eigvals.double().clone()
    scale = float(base.T.abs().max())
    del base
    torch.cuda.empty_cache()
    res = hlv.lanczos(op, m, v0, reorth="full", reorth_tol=1e-5)
    Q = res.Q
    G = torch.zeros(m, m, dtype=torch.float64, device=cuda_dev)
    for c0 in range(0, Q.shape[1], 1 << 22):
        Qc = Q[:, c0: c0 + (1 << 22)].double()
        G += Qc @ Qc.t()
    orth = float((G - torch.eye(m, dtype=torch.float64, device=cuda_dev)).abs().max())
    rec = {"max_abs_QQt_minus_I": orth, "passes_applied": res.conditional_passes, "of": m,
           "alpha_diff_vs_unconditional": float((res.alphas - a0).abs().max()) / scale,
           "beta_diff_vs_unconditional": float((res.betas - b0).abs().max()) / scale,
           "ritz_diff_vs_unconditional": float((res.eigvals.double() - ev0).abs().max()) / scale}
    _record("r02_reorth_tol_full_size_test.json", rec)
    print(json.dumps(rec))
    assert orth < 1e-5, rec
    assert rec["alpha_diff_vs_unconditional"] < 1e-5 and rec["beta_diff_vs_unconditional"] < 1e-5, rec
    assert rec["ritz_diff_vs_unconditional"] < 1e-5, rec
