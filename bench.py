#!/usr/bin/env python
"""Benchmark: Lanczos iterations/s for GPT-2 124M, m=100, full CGS2 reorthogonalisation, fp32 basis.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                          # the reference's CPU path (gpt2_hessian_cpu.py shape)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is ONE Lanczos iteration (lanczostrain_hand.py:188-203 order): one Hessian-vector
product of the GPT-2 loss over the global token batch by torch double-backward, the fused
gather + alpha, the three-term update, two passes of classical Gram-Schmidt against every stored
basis row, normalise + store.  The m=100 run has basis depth j = 0..99; with --steps 100 (default)
the timed region IS that run; with another K, step i runs at depth floor((i+0.5)*100/K) over a
pre-built orthonormal basis so the mean depth (hence mean cost) is that of the m=100 run.

The forward pass and the first backward of the double-backward do not depend on the Lanczos vector.
The headline arms (--hvp reuse, the library default) run them ONCE per Lanczos run -- inside the timed
region, at its first step -- and every iteration then performs the second backward (bit-identical
results, tests/test_gpu_zz_full_size.py); --hvp rebuild redoes them every iteration like the reference
(round 1's headline; reported under "extras" at N=1).

Scaling is STRONG: the global batch (8 sequences x 512 tokens, the batch of the reference's logged
runs) and therefore the operator and T are the same at every N; ranks shard the sequences
(reduce-scatter of Hv) and the basis along the parameter dimension (k-float all-reduces).
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# Rank 0 prints ONE JSON line on stdout and nothing else: keep a private handle to the real stdout and point
# fd 1 at stderr, so whatever a library writes to stdout (NCCL's version banner, warnings) lands in stderr.
_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


import torch  # noqa: E402

M_DEPTH = 100
SEQ_LEN = 512
VOCAB = 50257
METRIC = "lanczos_iters_per_sec_gpt2_124m_m100_cgs2"


# --------------------------------------------------------------------------- helpers
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--global-batch", type=int, default=8, help="sequences in the Hessian's batch (reference's logged runs: 8)")
    ap.add_argument("--micro-batch", type=int, default=0, help="sequences per double-backward; 0 = min(8, global_batch/ranks)")
    ap.add_argument("--prefill", default="lanczos", choices=["lanczos", "random"],
                    help="how the depth-100 basis is built before a sampled-depth run (random: no kernels; for short profiling runs)")
    ap.add_argument("--basis-dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="CGS2 as 4 separate passes (project, update, project, update)")
    ap.add_argument("--extras", action="store_true", help="also time the cached-first-backward HVP modes (reported under 'extras'); default at N=1")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--reorth-tol", type=float, default=None,
                    help="NOT the headline: apply the last Gram-Schmidt pass only when a projection coefficient exceeds tol*|w| "
                         "(gpytorch's rule, tol=1e-5), decided on the device; default = unconditional two-pass CGS")
    ap.add_argument("--blas", default="default", choices=["default", "cublas", "cublaslt"],
                    help="torch.backends.cuda.preferred_blas_library for the HVP's fp32 GEMMs (probe; fp32 either way)")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "on", "off"],
                    help="graph mode: capture the application as two graphs and prefetch the v-independent half of iteration j+1 on a "
                         "side stream while the recurrence and collectives of iteration j run.  auto = on for N>1 (+2%% at N=8), off at "
                         "N=1 (+0.5%% there, and the concurrent GEMMs would blur the per-kernel roofline timings)")
    ap.add_argument("--hvp", default="reuse", choices=["reuse", "rebuild"],
                    help="reuse (library default): forward + first backward once per Lanczos run, inside the timed region, then one second "
                         "backward per iteration; rebuild: the whole double-backward every iteration, like the reference")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1: how Hv shards, Lanczos vectors and Gram-Schmidt coefficients move between ranks.  peer = libhlv kernels over "
                         "NVLink peer memory (reduce-scatter + alpha, coefficient exchange in the kernels' epilogues/prologues, v written "
                         "straight into every peer); nccl = torch.distributed collectives; auto = peer when symmetric memory is available")
    ap.add_argument("--hvp-mode", default="graph", choices=["graph", "eager"],
                    help="graph: the whole double-backward (forward, both backward passes, gather) is captured once into a CUDA "
                         "graph and replayed every iteration -- all of the work, none of the ~4,000 Python-issued launches; "
                         "eager: issue it from Python every iteration like the reference")
    ap.add_argument("--workload", default="lanczos", choices=["lanczos", "ritz", "adjust"],
                    help="lanczos: the headline (default).  ritz: materialise all 100 Ritz vectors V = Y^T Q of the m=100 GPT-2 run "
                         "(gpt2_hessian_cpu.py:217); adjust: the low-rank gradient adjustment with k=100 pairs (vector_adjust.cu) -- "
                         "kernel-only legs with their own roofline, N=1")
    ap.add_argument("--small", action="store_true", help="tiny model for a functional check of this script (NOT a benchmark)")
    return ap.parse_args()


def depth_schedule(K: int):
    return [min(M_DEPTH - 1, int((i + 0.5) * M_DEPTH / K)) for i in range(K)]


def build_model(small: bool):
    from transformers import GPT2Config, GPT2LMHeadModel
    if small:
        cfg = GPT2Config(vocab_size=1024, n_positions=64, n_embd=64, n_layer=2, n_head=2, attn_implementation="eager",
                         resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    else:   # the reference's model: GPT2Config(vocab_size=len(tokenizer), n_positions=512)  gpt2_hessian_cpu.py:142-143
        cfg = GPT2Config(vocab_size=VOCAB, n_positions=SEQ_LEN, attn_implementation="eager")
    torch.manual_seed(0)
    return GPT2LMHeadModel(cfg).eval(), cfg


def make_tokens(cfg, global_batch: int, micro_batch: int, seq_len: int):
    g = torch.Generator().manual_seed(1234)            # precedent: gpt2_savehessian_noise.py:42-46
    ids = torch.randint(0, cfg.vocab_size, (global_batch, seq_len), generator=g)
    return [ids[i: i + micro_batch].contiguous() for i in range(0, global_batch, micro_batch)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel: str, algorithmic_bytes_per_launch: float):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu --set full
    capture (profiles/roofline_traffic.json): the captured launch's traffic/algorithmic ratio applied to this
    run's mean algorithmic bytes per launch (the basis depth differs from launch to launch)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p)).get(kernel)
            return algorithmic_bytes_per_launch * d["dram_bytes"] / d["algorithmic_bytes"]
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------- reference's CPU path (oracle port)
def host_threads() -> int:
    """All the host threads the box offers.  torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the
    reference arm runs on rank 0 alone, so it takes the whole machine back explicitly."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def reference_iteration_sampler(model_dev, batches_dev, weights, n, dev, host_rows: int):
    """The faithful gpt2_hessian_cpu.py shape: HVP on the GPU by the reference's own
    formulation (sum(v*g).backward(), torch.cat), `.cpu()` of the result every iteration
    (gpt2_hessian_cpu.py:137), recurrence + full reorthogonalisation on HOST cores with torch CPU ops,
    against a REAL host-resident basis of ``host_rows`` rows (49.6 GB at depth 100) -- nothing is extrapolated.
    Returns step(depth) -> seconds (wall, everything included)."""
    import oracle   # CPU baseline leg only (the checker, timed as the baseline -- never the product path)
    g = torch.Generator().manual_seed(99)
    Qh = torch.empty(host_rows, n)
    base = torch.randn(n, generator=g)
    base /= base.double().norm().float()
    for r in range(host_rows):           # synthetic unit-norm host rows (values do not affect timing; one threaded pass per row)
        torch.mul(base, 1.0 if r % 2 == 0 else -1.0, out=Qh[r])
        Qh[r, r::host_rows] *= 0.5       # rows differ, cheaply
    state = {"v": Qh[0].clone(), "v_old": Qh[min(1, host_rows - 1)].clone()}

    def step(depth_rows: int):
        t0 = time.perf_counter()
        v = state["v"]
        w = oracle.hess_vec_dataset(v.to(dev), batches_dev, model_dev, weights=weights)   # H2D of v + GPU HVP
        w = w.cpu()                                                                       # D2H, :137
        alpha = torch.dot(w, v)                                                           # lanczostrain_hand.py:200
        w -= (alpha * v + 0.5 * state["v_old"])                                           # :202
        rows = min(depth_rows, host_rows)
        for _ in range(2):                                                                # CGS2 on host cores
            c = Qh[:rows] @ w
            w -= Qh[:rows].t() @ c
        b = torch.norm(w, 2)                                                              # :190-193
        state["v_old"], state["v"] = v, w / b
        if rows < host_rows:
            Qh[rows].copy_(state["v"])                                                    # :194  Q[i+1] = v
        return time.perf_counter() - t0
    return step


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_threads()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, cfg = build_model(args.small)
    model.to(dev)
    seq = 64 if args.small else SEQ_LEN
    batches = [b.to(dev) for b in make_tokens(cfg, args.global_batch, args.micro_batch, seq)]
    weights = [b.shape[0] / args.global_batch for b in batches]
    n = sum(p.numel() for p in model.parameters())
    sched = depth_schedule(args.steps)
    host_rows = max(sched) + 1
    step = reference_iteration_sampler(model, batches, weights, n, dev, host_rows)
    for i in range(args.warmup):
        step(sched[i % len(sched)] + 1)
    clocks = ClockSampler(0); clocks.start()
    t_wall0 = time.perf_counter()
    for j in sched:
        step(j + 1)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    value = args.steps / wall
    sample = (f"{args.steps} iterations, every one measured in full: GPU HVP by the reference's formulation ({args.global_batch}x{seq} tokens) "
              f"+ .cpu() + host three-term update + host CGS2 against a real {host_rows}-row host basis ({host_rows * n * 4 / 1e9:.1f} GB) at the "
              f"scheduled depth (mean {sum(sched) / len(sched) + 1:.1f} rows) with {cores} host threads; wall {wall:.1f}s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, seq, n),
            "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample,
                             "os_cpu_count": os.cpu_count()},
            "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "clocks": clocks.stop(), "gpu_launches": 0}
    emit(line)


def bench_config(args, seq, n):
    return {"workload": "GPT-2 124M (GPT2Config(vocab_size=50257, n_positions=512), random init seed 0, eager attention, fp32, TF32 off) "
                        "Lanczos m=100 with full two-pass classical Gram-Schmidt reorthogonalisation, "
                        f"{'fp32' if args.basis_dtype == 'f32' else 'bf16'} basis" if not args.small else "SMALL functional check (not a benchmark)",
            "reorth_tol": args.reorth_tol, "P": n, "global_batch": args.global_batch, "micro_batch": args.micro_batch, "seq_len": seq, "basis_depth": M_DEPTH,
            "depth_schedule": "j=0..99 (the m=100 run itself)" if args.steps == M_DEPTH else f"floor((i+0.5)*100/{args.steps}) over a pre-built orthonormal basis",
            "l2": "no flush: every pass streams >= 0.5 GB per basis row, far beyond the 126 MB L2",
            "parallelism": f"{args.gpus} rank(s): micro-batches sharded (reduce-scatter of Hv), basis sharded along P (k-float all-reduce)"}


# --------------------------------------------------------------------------- this repo's arm
def run_ours(args, rank, world, local_rank):
    import hessian_llm_vision_b200 as hlv
    from hessian_llm_vision_b200 import kernels
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if args.blas != "default":
        torch.backends.cuda.preferred_blas_library(args.blas)
    comm = hlv.Comm()
    model, cfg = build_model(args.small)
    model.to(dev)
    seq = 64 if args.small else SEQ_LEN
    n = sum(p.numel() for p in model.parameters())
    n_micro = args.global_batch // args.micro_batch
    assert n_micro >= world and args.global_batch % args.micro_batch == 0, "need at least one micro-batch per rank"
    all_batches = make_tokens(cfg, args.global_batch, args.micro_batch, seq)
    mine_host = [b.pin_memory() for b in hlv.shard_batches(all_batches, rank, world)]
    mine_dev = [b.to(dev) for b in mine_host]
    reuse = args.hvp == "reuse"
    new_op = lambda batches, keep: hlv.HessianVectorProduct(model, batches, total_sequences=args.global_batch, cache_graph=keep, device=dev)
    op_dev, op_host = new_op(mine_dev, reuse), new_op(mine_host, reuse)
    hvp_modes = {}

    def graphed(op, tag, keep_first):
        """The operator the timed loop uses: the same double-backward, replayed from CUDA graphs (or eager)."""
        if args.hvp_mode != "graph":
            hvp_modes[tag] = "eager, " + ("first-backward graph kept for the run" if keep_first else "rebuilt every iteration")
            return op
        try:
            pipe = (not keep_first) and (args.pipeline == "on" or (args.pipeline == "auto" and world > 1))
            g = op.capture(out=eng.w if world == 1 else eng.hv_full, pipeline=pipe, reuse_first=keep_first)
            hvp_modes[tag] = ("cuda_graph: " + ("forward + first backward replayed ONCE per run inside the timed region, second backward + gather every iteration"
                                                if keep_first else "forward + both backward passes + gather replayed every iteration")
                              + ("; the v-independent half of iteration j+1 overlaps the recurrence of iteration j on a side stream" if pipe else ""))
            return g
        except Exception as e:  # noqa: BLE001  -- a model that cannot be captured still benches, eagerly
            hvp_modes[tag] = f"eager (capture failed: {repr(e)[:200]})"
            torch.cuda.synchronize()
            return op
    basis_dtype = torch.float32 if args.basis_dtype == "f32" else torch.bfloat16
    eng = hlv.LanczosEngine(op_dev, n, M_DEPTH, dev, reorth="full", basis_dtype=basis_dtype, comm=comm, profile=True,
                            fused_cgs=not args.no_fused, reorth_tol=args.reorth_tol, exchange=args.exchange)
    torch.manual_seed(7)                                  # probe: randn(P)/norm, diego_pythia.py:147-149
    v0 = torch.randn(n)
    v0 = v0.to(dev)                                       # normalise ON THE DEVICE: torch's CPU float32 norm of 1.24e8 elements
    v0 /= torch.linalg.vector_norm(v0.double()).float()   # is 1.4% off (profiles/README.md), which would leave |v0|^2 - 1 = 2.7e-2
    sched = depth_schedule(args.steps)
    real_run = args.steps == M_DEPTH

    def prefill():
        """Untimed setup: an orthonormal basis of depth 100 from a recurrence-only Lanczos run on a
        synthetic diagonal operator (same kernels, no HVP)."""
        g = torch.Generator(device=dev).manual_seed(11)
        diag = torch.randn(n, device=dev, generator=g) / world     # every rank contributes diag*v/world; the sum is diag*v
        keep, eng.hvp = eng.hvp, (lambda v: diag * v)
        eng.start(v0)
        for j in range(M_DEPTH):
            eng.step(j)
        eng.hvp = keep

    def timed(op, e2e: bool):
        eng.hvp = op
        if real_run:
            eng.start(v0)                                 # also invalidates a captured operator's first half
        elif hasattr(op, "invalidate"):
            op.invalidate()                               # the run's forward + first backward belong to the timed region
        elif hasattr(op, "clear_cache"):
            op.clear_cache()
        eng.phases.pairs.clear()
        comm.barrier(); torch.cuda.synchronize()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        launches0 = kernels.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if os.environ.get("HLV_PROFILE_RANGE") == "1" and not e2e:     # `ncu --profile-from-start off`: only the timed steps
            torch.cuda.profiler.start()
        ev0.record()
        sink = 0.0
        for j in sched:
            eng.step(j)
            if e2e:                                       # device -> host read of the step's result (alpha_j, beta_{j+1})
                sink += float(eng.alphas[j].item()) + float(eng.betas[j + 1].item())
        if hasattr(op, "drain"):
            op.drain()                                    # a prefetched half-application is work of this region: wait for it
        ev1.record()
        torch.cuda.synchronize(); comm.barrier()
        if os.environ.get("HLV_PROFILE_RANGE") == "1" and not e2e:
            torch.cuda.profiler.stop()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()), kernels.launch_count - launches0, (clocks.stop() if rank == 0 else None), eng.phases.summary()

    if args.prefill == "random" and not real_run:
        g = torch.Generator(device=dev).manual_seed(11)
        for r in range(M_DEPTH):                           # timing does not depend on the values
            eng.basis[r].copy_(torch.randn(eng.basis.shape[1], device=dev, generator=g) * (1.0 / eng.basis.shape[1] ** 0.5))
        if world > 1:
            eng.v_full.normal_(generator=g).mul_(1.0 / n ** 0.5)
    else:
        prefill()
    run_dev = graphed(op_dev, "value", reuse)
    eng.hvp = run_dev
    for i in range(max(args.warmup, 0)):
        eng.step(sched[i % len(sched)])
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(dev)

    first0 = getattr(run_dev, "first_replays", None)
    ms, launches, clocks, phases = timed(run_dev, e2e=False)
    value = args.steps / (ms / 1e3)
    peak_bytes = torch.cuda.max_memory_allocated(dev)
    first_half_runs = (run_dev.first_replays - first0) if first0 is not None else None
    ritz_top = None
    if real_run:
        res = eng.result()
        ritz_top = [float(x) for x in res.eigvals[-3:]]

    e2e = None
    del run_dev
    eng.hvp = op_dev
    op_dev.clear_cache()
    torch.cuda.empty_cache()
    if not args.no_e2e:
        run_host = graphed(op_host, "e2e", reuse)
        eng.hvp = run_host
        for i in range(min(max(args.warmup, 0), 1)):
            eng.step(sched[i % len(sched)])
        h2d0 = op_host.h2d_bytes
        ms_e, _, _, _ = timed(run_host, e2e=True)
        h2d = op_host.h2d_bytes - h2d0
        del run_host
        eng.hvp = op_dev
        op_host.clear_cache()
        torch.cuda.empty_cache()
        e2e = {"value": args.steps / (ms_e / 1e3), "unit": "iterations/s",
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": 16,
               "h2d_bytes_total": h2d, "ms_per_step": ms_e / args.steps,
               "api": "LanczosEngine.step over HessianVectorProduct with pinned-host token batches; alpha/beta read back every step.  "
                      + ("The tokens feed only the forward pass, which --hvp reuse runs once per Lanczos run (inside the timed region): they are "
                         "copied H2D when it runs, h2d_bytes_per_step = that total / steps" if reuse else
                         "The tokens are copied H2D inside every application"),
               "hvp_mode": hvp_modes.get("e2e")}

    # ---- extras (NOT the headline) ---------------------------------------------------------------------------
    # The other HVP policy, same metric, same steps: with --hvp reuse (default) this is round 1's headline arm, which
    # redoes forward + first backward every iteration like the reference; plus the eager (no CUDA graph) variant.
    extras = {}
    if (args.extras or world == 1) and not args.no_extras:
        try:
            other = not reuse
            op_x = new_op(mine_dev, other)
            gop = graphed(op_x, "extra", other)
            eng.hvp = gop
            eng.step(sched[0])
            ms_x, _, _, _ = timed(gop, e2e=False)
            extras["hvp_reuse_first_backward" if other else "hvp_rebuilt_every_iteration"] = {
                "value": args.steps / (ms_x / 1e3), "ms_per_step": ms_x / args.steps, "hvp_mode": hvp_modes.get("extra")}
            del gop
            op_x.clear_cache()
            torch.cuda.empty_cache()
            del op_x
            eng.hvp = op_dev
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extras["error"] = repr(e)[:300]

    if eng.peer is not None and eng.peer.error():
        raise RuntimeError(f"peer exchange: a wait on channel {eng.peer.error() - 1} ran into its time limit; the numbers are invalid")
    if rank != 0:
        return
    # ---- roofline of the dominant libhlv kernel, from CUDA events inside the timed region ----
    peak, peak_src = measured_peak()
    s = 4 if args.basis_dtype == "f32" else 2
    n_loc = eng.shard_n if world > 1 else n
    G = world
    kern_bytes = {
        "cgs_project": lambda d: (d["rows"] * s + 4 * d["calls"]) * n_loc,
        "update_project": lambda d: (d["rows"] * s + 8 * d["calls"]) * n_loc,     # three-term update folded into the first projection: w read + written
        "cgs_update": lambda d: (d["rows"] * s + 8 * d["calls"]) * n_loc,
        "cgs_update_project": lambda d: (d["rows"] * s + 8 * d["calls"]) * n_loc,   # V once; w read + written
        "update": lambda d: 16 * n_loc * d["calls"],
        "normalize": lambda d: ((4 + s) if s == 4 else (4 + 4 + s)) * n_loc * d["calls"] + (4 * n_loc * (1 if eng.multicast else G - 1) * d["calls"] if eng.peer else 0),
        "gather": lambda d: (8 + 4) * n * d["calls"],      # read pieces 4n + write w 4n (+4n v for the fused alpha on the last micro-batch)
        "dot": lambda d: 8 * n_loc * d["calls"],
        "reduce_scatter_alpha": lambda d: (4 * (1 if eng.multicast else G) + 4 + 4) * n_loc * d["calls"],   # G shards in (G-1 over NVLink; one in-switch sum with multicast), w out, v in
    }
    kernels_out = {}
    for name, fn in kern_bytes.items():
        if name in phases and phases[name]["ms"] > 0:
            d = phases[name]
            gbs = fn(d) / (d["ms"] * 1e-3) / 1e9
            kernels_out[name] = {"ms_total": round(d["ms"], 3), "launches": d["calls"], "achieved_gbs": round(gbs, 1),
                                 "frac_of_peak": round(gbs / peak, 4), "bytes_per_launch": fn(d) / d["calls"]}
    if args.reorth_tol is not None and "cgs_update" in kernels_out:
        kernels_out["cgs_update"]["note"] = "predicated pass: bytes are counted as if every launch ran; skipped launches move none"
    ours_ms = sum(v["ms_total"] for v in kernels_out.values())
    ours_bytes = sum(v["bytes_per_launch"] * v["launches"] for v in kernels_out.values())
    bound_ms = ours_bytes / (peak * 1e9) * 1e3 / args.steps          # per step, this rank's shard, at the measured HBM peak
    hbm_roofline = {"recurrence_bytes_per_step_per_gpu": ours_bytes / args.steps, "bound_ms_per_step": bound_ms,
                    "iterations_per_s_at_roofline": 1e3 / bound_ms if bound_ms else None,
                    "recurrence_frac_of_roofline": bound_ms / (ours_ms / args.steps) if ours_ms else None,
                    "value_frac_of_roofline": (value * bound_ms / 1e3) if bound_ms else None,
                    "note": "recurrence kernels only (the libhlv launches timed by CUDA events; a gather captured inside the HVP graph is "
                            "not listed); the step as a whole is bound by the torch HVP, see hvp_ms_per_step"}
    top = max((k for k in kernels_out if k.startswith("cgs") or k == "update_project"), key=lambda k: kernels_out[k]["ms_total"], default=None)
    roofline = None
    if top:
        cname = {"update_project": "cgs_project"}.get(top, top)
        kname = f"hlv::{cname}_kernel<{'float' if s == 4 else 'bf16'}>"
        roofline = {"bound": "hbm", "kernel": kname, "achieved": kernels_out[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels_out[top]["frac_of_peak"], "bytes_per_launch": kernels_out[top]["bytes_per_launch"],
                    "traffic": ncu_traffic(top, kernels_out[top]["bytes_per_launch"]), "peak_source": peak_src,
                    "bytes_model": "project: (rows*s+4)*n (+4n when it also applies the three-term update); update and fused update_project: "
                                   "(rows*s+8)*n per launch (DESIGN.md section 3)",
                    "share_of_step": round(kernels_out[top]["ms_total"] / ms, 4)}
    line = {"metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(args, seq, n),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "kernels": kernels_out, "hbm_roofline": hbm_roofline,
            "recurrence_only": {"ms_per_step": ours_ms / args.steps, "iterations_per_s": args.steps / (ours_ms / 1e3) if ours_ms else None,
                                "note": "sum of libhlv kernel time (CUDA events) in the timed region; HVP (torch) excluded"},
            "hvp_ms_per_step": phases.get("hvp", {}).get("ms", 0.0) / args.steps,
            "phases_ms_per_step": {k: round(v["ms"] / args.steps, 4) for k, v in phases.items()},
            "hvp_policy": {"mode": args.hvp, "first_half_runs_in_timed_region": first_half_runs,
                           "peak_device_bytes": peak_bytes},
            "exchange": eng.exchange_mode,
            "hvp_mode": hvp_modes.get("value"), "ritz_top3": ritz_top, "extras": extras}
    # ---- CPU baseline: the reference's CPU path on this box's host cores (rank 0, N=1 only), a bounded sample ----
    if world == 1 and not args.no_cpu_baseline:
        del eng
        torch.cuda.empty_cache()
        cores = host_threads()
        k_s = 6                                           # 6 iterations spread over the m=100 depth range (~10-15 s of host work)
        sched_s = depth_schedule(k_s)
        host_rows = max(sched_s) + 1
        step = reference_iteration_sampler(model, mine_dev, op_dev.weights, n, dev, host_rows)
        step(sched_s[0] + 1)                              # warm
        t0 = time.perf_counter()
        for j in sched_s:
            step(j + 1)
        t_s = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": k_s / t_s, "unit": "iterations/s", "cores": cores, "kind": "port",
                                "sample": f"{k_s} iterations of the gpt2_hessian_cpu.py shape measured in full (GPU HVP + .cpu() + host recurrence + host CGS2 "
                                          f"against a real {host_rows}-row host basis) at depths {[j + 1 for j in sched_s]} (mean {sum(sched_s) / k_s + 1:.1f} rows, "
                                          f"the m=100 run's mean is 50.5); {t_s:.1f}s",
                                "os_cpu_count": os.cpu_count()}
    emit(line)


# --------------------------------------------------------------------------- kernel-only legs: Ritz vectors, adjustment
def run_aux(args):
    """Consumers of the basis after the run, at GPT-2 size, device-resident, CUDA events."""
    import numpy as np
    from hessian_llm_vision_b200 import kernels
    from hessian_llm_vision_b200.adjust import adjust_gradient_implicit
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    n = 1 << 20 if args.small else 124_046_592
    m = M_DEPTH
    g = torch.Generator(device=dev).manual_seed(3)
    Q = torch.empty(m, n, device=dev)
    for r in range(m):
        Q[r].normal_(generator=g).mul_(1.0 / n ** 0.5)
    Yt, _ = torch.linalg.qr(torch.randn(m, m, device=dev, generator=g, dtype=torch.float64))
    peak, peak_src = measured_peak()
    clocks = ClockSampler(0)
    reps, warm = max(args.steps if args.steps != M_DEPTH else 5, 1), max(args.warmup, 3)

    def time_it(fn):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.start()
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps, clocks.stop()
    if args.workload == "ritz":
        Y = Yt.float().contiguous()
        out = torch.empty(m, n, device=dev)
        l0 = kernels.launch_count
        ms, ck = time_it(lambda: kernels.ritz_vectors(Q, m, Y, out, n))
        launches = (kernels.launch_count - l0) // (reps + warm)
        nb = (m * 4 + m * 4) * n                              # read Q once, write V once
        ref = (Yt[:, :3].t() @ Q[:, : 1 << 16].double())
        err = float((out[:3, : 1 << 16].double() - ref).abs().max() / ref.abs().max())
        gbs = nb / (ms * 1e-3) / 1e9
        line = {"metric": "ritz_vectors_all_m100_gpt2_124m_ms", "value": ms, "unit": "ms", "higher_is_better": False,
                "roofline": {"bound": "hbm", "kernel": "hlv::ritz_vectors_tc_kernel (tcgen05 kind::tf32, 3xTF32)" if os.environ.get("HLV_RITZ_TC", "1") != "0"
                             else "hlv::ritz_vectors_kernel<float> (CUDA cores, 8 vectors per pass)",
                             "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "bytes_per_launch": nb, "traffic": None,
                             "peak_source": peak_src, "bytes_model": "(m*4 + nvec*4)*n: Q read once, V written once",
                             "tflops_fp32_equivalent": 2.0 * m * m * n / (ms * 1e-3) / 1e12},
                "max_rel_err_vs_float64_sample": err}
    else:
        k = m
        lam = torch.linspace(0.5, 20.0, k, device=dev)
        gvec = torch.randn(n, device=dev, generator=g)
        ws = kernels.Workspace(dev, max_rows=k + 1)
        outv = gvec.clone()
        l0 = kernels.launch_count
        ms, ck = time_it(lambda: adjust_gradient_implicit(gvec, Q, m, Yt.cpu().numpy(), lam.cpu().numpy(), 0.1, ws=ws, out=outv))
        launches = (kernels.launch_count - l0) // (reps + warm)
        nb = (2 * k * 4 + 16) * n                              # project (k rows + g) and update (k rows + out read/write)
        gbs = nb / (ms * 1e-3) / 1e9
        line = {"metric": "lowrank_adjust_k100_gpt2_124m_ms", "value": ms, "unit": "ms", "higher_is_better": False,
                "roofline": {"bound": "hbm", "kernel": "hlv::cgs_project_kernel<float> + hlv::cgs_update_kernel<float>", "achieved": gbs,
                             "peak": peak, "unit": "GB/s", "frac": gbs / peak, "bytes_per_launch": nb / 2, "traffic": None, "peak_source": peak_src,
                             "bytes_model": "(2*k*4 + 16)*n over the two passes (implicit form: g += Q^T (Y diag(s) Y^T) (Q g), no Ritz vectors formed)"}}
    line.update({"n_gpus": 1, "steps": reps, "warmup": warm, "ms_per_step": ms, "scaling": "n/a", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                 "config": {"workload": f"{args.workload}: m = k = {m}, P = {n}, fp32 basis resident in HBM", "l2": "no flush: every pass streams >= 49 GB"},
                 "gpu_launches": launches, "clocks": ck, "e2e": None})
    emit(line)


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.micro_batch <= 0:               # same rule in both arms, from --gpus, so that the two lines carry the same config
        args.micro_batch = max(1, min(8, args.global_batch // max(args.gpus, 1)))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "lanczos":
        if rank == 0:
            run_aux(args)
        return
    if world > 1:
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
