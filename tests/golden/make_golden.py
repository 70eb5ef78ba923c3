#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S
OWN SOURCE in the build container (it is read from /root/reference at
generation time only; nothing is copied into this repo except the numeric
outputs).  Re-run with:  python tests/golden/make_golden.py

What is executed, verbatim, from the reference tree:
  G1  Lanczos_Scratch/Discrepancy.ipynb cell 0 (seeded M, v) and the hand
      full-reorth loop of cell 1 (source lines from ``T = torch.zeros`` on; the
      gpytorch import/call above it cannot run here -- gpytorch is absent).
  G2  lanczostrain_hand.py lines 171-203 (the hand Lanczos loop), with
      ``hess_vec`` bound to a dense symmetric matvec.
  G3  the ``hess_vec`` function object of gpt2_hessian_cpu.py:75-109, extracted
      with ``ast`` and called with cuda=False on a tiny random-init GPT-2.
  G4  the ``hess_vec`` function object of gpt2_savehessian.py:130-163
      (dataset loop), same tiny model, two batches.
      NOTE: that function reads a module global ``num_gpus`` and scales by
      ``len(batch)/N`` where batch is a dict (quirk Q6) -- the raw output is
      stored together with the scale so the oracle can be compared modulo Q6.
  G5  the shipped result dicts (eigeninfo/*/results.ckpt) -> layout fixture.
"""
from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _func_source(path: str, name: str) -> str:
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def g1_discrepancy():
    nb = json.load(open(f"{REF}/Lanczos_Scratch/Discrepancy.ipynb"))
    cell0 = "".join(nb["cells"][0]["source"])
    cell1 = "".join(nb["cells"][1]["source"])
    loop_src = cell1[cell1.index("T = torch.zeros([lanczos_iters, lanczos_iters])"):]
    out = {}
    for m in (2, 4, 8, 16):
        ns = {"torch": torch}
        exec(cell0, ns)                         # seeds 42, builds M, v
        ns["lanczos_iters"] = m
        ns["Hess_Vec_Orig"] = lambda M, v: torch.matmul(M, v)   # cell 1:5-7 (one line body)
        exec(loop_src, ns)
        out[f"T_m{m}"] = ns["T"].numpy().copy()
        out[f"Q_m{m}_head"] = torch.stack(ns["u_list"])[:, :8].numpy().copy()
    np.savez(f"{OUT}/discrepancy_reorth.npz", **out)
    print("G1 T(m=2) =", out["T_m2"].tolist())


def g2_hand_loop():
    lines = open(f"{REF}/lanczostrain_hand.py").read().split("\n")
    body = "\n".join(l[8:] if l.startswith("        ") else l for l in lines[170:203])  # 171..203, de-indented
    out = {}
    for seed, n, k in ((42, 1000, 3), (7, 512, 10), (3, 2048, 24)):
        torch.manual_seed(seed)
        M = torch.randn(n, n)
        M = (M + M.t()) / 2
        v = torch.randn(n)
        random_vec = v / torch.norm(v, 2)
        ns = {
            "torch": torch, "device": "cpu", "P": n, "lanczos_iters": k,
            "random_vec": random_vec.clone(), "input_ids": None, "model": None,
            "hess_vec": (lambda vec, ids, model, device, M=M: torch.matmul(M, vec)),
            "time": (lambda: 0.0),
        }
        exec(body, ns)
        out[f"T_s{seed}_n{n}_k{k}"] = ns["T"].numpy().copy()
        out[f"Qhead_s{seed}_n{n}_k{k}"] = ns["Q"][:, :8].numpy().copy()
    np.savez(f"{OUT}/hand_lanczos.npz", **out)
    print("G2 T diag (seed42,k3) =", np.diag(out["T_s42_n1000_k3"]).tolist())


def _tiny_gpt2():
    from transformers import GPT2Config, GPT2LMHeadModel
    cfg = GPT2Config(vocab_size=97, n_positions=16, n_embd=16, n_layer=2, n_head=2,
                     attn_implementation="eager", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    torch.manual_seed(0)
    return GPT2LMHeadModel(cfg)


def g3_g4_hvp():
    model = _tiny_gpt2()
    state = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    P = sum(p.numel() for p in model.parameters())
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, 97, (4, 16), generator=g)
    ids2 = torch.randint(0, 97, (2, 16), generator=g)
    torch.manual_seed(5)
    vec = torch.randn(P)
    vec /= vec.norm()

    # G3: single-batch hess_vec, gpt2_hessian_cpu.py:75-109
    ns = {"torch": torch}
    exec(_func_source(f"{REF}/gpt2_hessian_cpu.py", "_bn_train_mode"), ns)
    exec(_func_source(f"{REF}/gpt2_hessian_cpu.py", "hess_vec"), ns)
    hv = ns["hess_vec"](vec, ids, model, cuda=False, bn_train_mode=False).detach().clone()

    # G4: dataset hess_vec, gpt2_savehessian.py:130-163 (needs num_gpus global, dict batches,
    # hard-coded .to("cuda") on the ids -> patch the device string only)
    src = _func_source(f"{REF}/gpt2_savehessian.py", "hess_vec").replace('.to("cuda")', '.to("cpu")')
    ns2 = {"torch": torch, "num_gpus": 1}
    exec(_func_source(f"{REF}/gpt2_savehessian.py", "_bn_train_mode"), ns2)
    exec(src, ns2)

    class _Loader(list):
        pass
    loader = _Loader([{"input_ids": ids, "attention_mask": torch.ones_like(ids)},
                      {"input_ids": ids2, "attention_mask": torch.ones_like(ids2)}])
    loader.dataset = list(range(6))            # N = 6 sequences
    hv_ds = ns2["hess_vec"](vec, loader, model, cuda=False, bn_train_mode=False).detach().clone()

    np.savez(f"{OUT}/tiny_gpt2_hvp.npz", ids=ids.numpy(), ids2=ids2.numpy(), vec=vec.numpy(),
             hv=hv.numpy(), hv_dataset_q6=hv_ds.numpy(), q6_scale=np.float64(2.0 / 6.0),
             **{"state." + k: v for k, v in state.items()})
    print("G3 P =", P, " |Hv| =", float(hv.norm()), " G4 |Hv_ds| =", float(hv_ds.norm()))


def g5_result_dicts():
    out = {}
    for tag, path in (
        ("m25", f"{REF}/eigeninfo/gpt2_subsample=0.0001_iters=25_basis=False/results.ckpt"),
        ("m35", f"{REF}/eigeninfo/gpt2_subsample=0.0001_iters=35_basis=False/results.ckpt"),
        ("m30", f"{REF}/Lanczos_Scratch/model_trained.pt.ckpt"),
    ):
        d = torch.load(path, weights_only=True, map_location="cpu")
        out[f"{tag}_keys"] = np.array(sorted(d.keys()))
        out[f"{tag}_eigvals"] = d["eigvals"].numpy()
        out[f"{tag}_gammas"] = d["gammas"].numpy()
    np.savez(f"{OUT}/shipped_result_dicts.npz", **out)
    print("G5 keys:", out["m25_keys"].tolist(), "m =", [out[f"{t}_eigvals"].shape[0] for t in ("m25", "m35", "m30")])


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    torch.set_num_threads(1)                   # single-thread: reduction order independent of core count
    g1_discrepancy()
    g2_hand_loop()
    g3_g4_hvp()
    g5_result_dicts()
