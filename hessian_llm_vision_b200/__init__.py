"""hessian_llm_vision_b200: B200-native Lanczos / stochastic-Lanczos-quadrature engine.

The package directory carries an importable name; ``hessian-llm-vision_b200`` at the repo root is a
symlink to it (the project's spelling).  The compute path is libhlv.so (hand-written sm_100a
CUDA behind the C ABI of include/hlv.h); there is no CPU or pure-PyTorch fallback.
"""
from .lanczos import Comm, LanczosEngine, LanczosResult, lanczos, lanczos_tridiag  # noqa: F401
from .hvp import (CurvVecProduct, HessianVectorProduct, criterion_loss, lm_loss,  # noqa: F401
                  shard_batches)
from .ritz import dense_T, ritz_values, slq_density, tridiag_eigh  # noqa: F401
from .results import (eigeninfo_path, load_eigeninfo, save_eigeninfo,  # noqa: F401
                      save_tridiagonal_checkpoint)

from .spectra import SLQResult, per_block_spectra, probe_vector, slq  # noqa: F401

__version__ = "0.2.0"


def library_path() -> str:
    from . import _lib
    return _lib.LIB_PATH
