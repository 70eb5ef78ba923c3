"""BASELINE.json configurations at FULL size on synthetic data (the bench line is config 2; these are the
others).  Each prints one JSON object: sizes, device time, iterations/s, and the size-independent checks the
parity tests use at small size (orthogonality of the stored rows, sum(gammas) = 1, sum(gammas*eigvals) = alpha_0).

  python scripts/run_configs.py --config 1            # GPT-2, m=25, no reorth, 20 sequences streamed as micro-batches
  python scripts/run_configs.py --config 3 [--iters 20]   # one Lanczos run per transformer block (12 runs, P=7,087,872)
  torchrun --nproc-per-node N scripts/run_configs.py --config 4   # Pythia-1.4B shapes, m=50, bf16 basis sharded along P
  [torchrun --nproc-per-node N] python scripts/run_configs.py --config 5 [--probes 16]   # SLQ: ResNet-50/10-class 16 probes x 80 + GPT-2 probes x 80; probes dealt over the ranks

Reference call sites: gpt2_hessian_cpu.py:207-216 / gpt2_savehessian.py:143-163 (1), ipynbs/visual-eigen.ipynb
cells 10-12 (3), diego_pythia.py:95-192 (4), d.sh:4-11 + train_savespec.py:61-91 (5)."""
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


def gpt2(dev):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(0)
    cfg = GPT2Config(vocab_size=50257, n_positions=512, attn_implementation="eager")
    return GPT2LMHeadModel(cfg).eval().to(dev), cfg


def tokens(vocab, n_seq, seq_len, micro, seed=1234):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (n_seq, seq_len), generator=g)
    return [ids[i: i + micro].contiguous() for i in range(0, n_seq, micro)]


def timed_run(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / 1e3, time.perf_counter() - t0


def checks(res, comm=None):
    """Size-independent properties of a finished run."""
    out = {"sum_gammas": float(res.gammas.sum()),
           "sum_gammas_eigvals_minus_alpha0": float((res.gammas.double() * res.eigvals.double()).sum() - res.alphas[0]),
           "ritz_min": float(res.eigvals[0]), "ritz_max": float(res.eigvals[-1]), "m": res.m, "breakdown": res.breakdown}
    if res.basis is not None:
        Q = res.Q                                                # [m, n_local]; Gram in column chunks (a bf16 shard of
        G = torch.zeros(res.m, res.m, dtype=torch.float64, device=Q.device)     # Pythia-1.4B is 70 GB: never upcast it whole)
        for c0 in range(0, Q.shape[1], 1 << 22):
            Qc = Q[:, c0: c0 + (1 << 22)].float()
            G += (Qc @ Qc.t()).double()
        if comm is not None and comm.world > 1:
            comm.all_reduce_sum(G)
        out["max_abs_QQt_minus_I"] = float((G - torch.eye(res.m, dtype=torch.float64, device=G.device)).abs().max())
    return out


def config1(args, dev):
    model, cfg = gpt2(dev)
    batches = [b.to(dev) for b in tokens(cfg.vocab_size, 20, 512, 8)]       # 20 = int(1e-4 * 205,328) documents
    op = hlv.HessianVectorProduct(model, batches)                            # B_i/N weights: the dataset mean
    m = args.iters or 25
    v0 = hlv.probe_vector(op.n, 0, dev)
    res, t_dev, _ = timed_run(lambda: hlv.lanczos(op, m, v0, reorth=None, keep_basis=True))
    return {"config": "1: GPT-2 124M, Lanczos 25 iters, no reorth, 20 sequences (micro-batches 8+8+4)", "P": op.n, "iters": m,
            "seconds": t_dev, "iterations_per_s": m / t_dev, **checks(res),
            "note": "no reorthogonalisation: the stored rows lose orthogonality by design (SURVEY F4); sum(gammas)=1 still holds"}


def config3(args, dev):
    model, cfg = gpt2(dev)
    batches = [b.to(dev) for b in tokens(cfg.vocab_size, 8, 512, 8)]
    m = args.iters or 20
    (evs, gms), t_dev, _ = timed_run(lambda: hlv.per_block_spectra(model, batches, m, seed=0, reorth="full"))
    P_blk = sum(p.numel() for p in model.transformer.h[0].parameters())
    return {"config": "3: GPT-2 per-transformer-block spectra, one Lanczos run per block", "blocks": len(evs), "P_block": P_blk,
            "iters_per_block": m, "seconds": t_dev, "iterations_per_s": len(evs) * m / t_dev,
            "ritz_max_per_block": [round(float(e[-1]), 5) for e in evs],
            "sum_gammas_per_block": [round(float(g.sum()), 6) for g in gms]}


def config4(args, dev, comm):
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM
    cfg = GPTNeoXConfig(vocab_size=50304, hidden_size=2048, num_hidden_layers=24, num_attention_heads=16, intermediate_size=8192,
                        max_position_embeddings=2048, tie_word_embeddings=False, attn_implementation="eager",
                        hidden_dropout=0.0, attention_dropout=0.0)
    torch.manual_seed(0)
    with torch.device(dev):
        model = GPTNeoXForCausalLM(cfg).eval()
    per_rank = args.seqs_per_rank
    n_seq = per_rank * comm.world
    mine = [b.to(dev) for b in hlv.shard_batches(tokens(cfg.vocab_size, n_seq, 512, per_rank), comm.rank, comm.world)]

    def loss(model, batch):                                                  # diego_pythia.py:105-107
        return model(input_ids=batch, labels=batch, use_cache=False).loss
    op = hlv.HessianVectorProduct(model, mine, loss_fn=loss, total_sequences=n_seq)
    m = args.iters or 50
    v0 = hlv.probe_vector(op.n, 0, dev)                                      # same seed on every rank
    torch.cuda.reset_peak_memory_stats(dev)
    res, t_dev, _ = timed_run(lambda: hlv.lanczos(op, m, v0, reorth="full", basis_dtype=torch.bfloat16, comm=comm))
    return {"config": f"4: Pythia-1.4B shapes, Lanczos {m} iters, bf16 basis sharded along P over {comm.world} GPU(s)", "P": op.n,
            "world": comm.world, "global_sequences": n_seq, "iters": m, "seconds": t_dev, "iterations_per_s": m / t_dev,
            "basis_rows_dtype": str(res.basis.dtype), "basis_shard_shape": list(res.basis.shape),
            "peak_mem_gb_rank0": torch.cuda.max_memory_allocated(dev) / 1e9, **checks(res, comm),
            "note": "bf16 storage of the rows: orthogonality ~1e-3 (one bf16 ulp), by design"}


def config5(args, dev, comm):
    import torch.nn as nn
    import torchvision
    rep = comm if comm.world > 1 else None          # probes dealt round-robin over the ranks, replicas only (d.sh:4-11 runs them one by one)
    out = {"config": "5: stochastic Lanczos quadrature, 80 iterations per probe", "ranks": comm.world,
           "probe_split": "round-robin over ranks, no data-path collective; (eigvals, gammas) exchanged once at the end"}
    torch.manual_seed(0)
    net = torchvision.models.resnet50(num_classes=10).to(dev)
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(128, 3, 32, 32, generator=g).to(dev), torch.randint(0, 10, (128,), generator=g).to(dev)
    op = hlv.HessianVectorProduct(net, [(x, y)], loss_fn=hlv.criterion_loss(nn.CrossEntropyLoss()), bn_train_mode=True)
    seeds = list(range(16))
    r, t_dev, _ = timed_run(lambda: hlv.slq(op, op.n, 80, seeds, dev, replicas=rep))
    grid, dens = r.density(num_points=512)
    out["resnet50"] = {"P": op.n, "probes": len(seeds), "iters": 80, "seconds": t_dev, "iterations_per_s": len(seeds) * 80 / t_dev,
                       "sum_gammas_eigeninfo": float(r.eigeninfo()["gammas"].sum()), "density_integral": float(((dens[1:] + dens[:-1]) * 0.5 * (grid[1:] - grid[:-1])).sum()),
                       "ritz_max_per_probe": [round(float(e[-1]), 4) for e in r.eigvals]}
    del op, net
    torch.cuda.empty_cache()
    model, cfg = gpt2(dev)
    batches = [b.to(dev) for b in tokens(cfg.vocab_size, 8, 512, 8)]
    op2 = hlv.HessianVectorProduct(model, batches).capture()
    seeds2 = list(range(args.probes))
    r2, t_dev2, _ = timed_run(lambda: hlv.slq(op2, op2.n, 80, seeds2, dev, replicas=rep))
    out["gpt2"] = {"P": op2.n, "probes": len(seeds2), "of": 16, "iters": 80, "seconds": t_dev2, "iterations_per_s": len(seeds2) * 80 / t_dev2,
                   "sum_gammas_eigeninfo": float(r2.eigeninfo()["gammas"].sum()),
                   "ritz_max_per_probe": [round(float(e[-1]), 4) for e in r2.eigvals]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[1, 3, 4, 5])
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--probes", type=int, default=2, help="config 5, GPT-2 leg: probes actually run (of the 16)")
    ap.add_argument("--seqs-per-rank", type=int, default=2, help="config 4: sequences of 512 tokens per rank")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    comm = hlv.Comm()
    try:
        if args.config == 1:
            out = config1(args, dev)
        elif args.config == 3:
            out = config3(args, dev)
        elif args.config == 4:
            out = config4(args, dev, comm)
        else:
            out = config5(args, dev, comm)
        if comm.rank == 0:
            print(json.dumps(out), flush=True)
    finally:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
