"""The reference's on-disk result layout (``eigeninfo`` / ``lanczos_results``).

  spectrum dict ... {'eigvals': f32[m] ascending, 'gammas': f32[m][, 'V': f32[m,P]]}
                    torch.save'd to <ckpt dir>/subsample={s}_iters={m}_basis={b}[suffix]/<ckpt name>.ckpt
                    (gpt2_savehessian.py:216-236; suffix '_noise' gpt2_savehessian_noise.py,
                    '_layeronly' gpt2_savehessian_layer.py; 'V' only in train_savespec.py:328-338)
  Pythia T ........ dense (k+1)x(k+1) T torch.save'd every iteration to
                    <dir>/diego_data_seed={d}_vector_seed={v}/ckpt.pt   (diego_pythia.py:127-130,192)
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch


def eigeninfo_path(checkpoint: str, subsample, iters: int, basis: bool, suffix: str = "") -> str:
    """Path rule of gpt2_savehessian.py:225-236: folder next to the checkpoint, file named after it."""
    parts = checkpoint.split("/")
    save_folder = "/".join(parts[:-1])
    save_name = parts[-1]
    folder = "subsample={}_iters={}_basis={}{}".format(str(subsample), str(iters), str(basis), suffix)
    return os.path.join(save_folder, folder, save_name + ".ckpt")


def save_eigeninfo(result, path: str, basis: bool = False) -> Dict[str, torch.Tensor]:
    """Write ``{'eigvals','gammas'[,'V']}`` (CPU float32 tensors) to ``path``; returns the dict.
    ``result`` is a LanczosResult (or an already-built dict)."""
    d = result if isinstance(result, dict) else result.eigeninfo(basis=basis)
    d = {k: v.detach().to("cpu", torch.float32) for k, v in d.items()}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(d, path)
    return d


def load_eigeninfo(path: str) -> Dict[str, torch.Tensor]:
    return torch.load(path, weights_only=True, map_location="cpu")


def save_tridiagonal_checkpoint(T: torch.Tensor, data_seed: int, vector_seed: int,
                                checkpoint_dir: str = "70mpythia") -> str:
    """diego_pythia.py:127-130 -- the per-iteration T checkpoint of the Pythia scripts."""
    d = os.path.join(checkpoint_dir, f"diego_data_seed={data_seed}_vector_seed={vector_seed}")
    os.makedirs(d, exist_ok=True)
    fn = os.path.join(d, "ckpt.pt")
    torch.save(T, fn)
    return fn
