"""TEST HELPER.  Several ranks of the peer exchange emulated on ONE GPU.

Every "rank" gets its own exchange area / Hv / v buffers on the same device and a PeerContext whose pointer tables
name the other ranks' buffers -- exactly what CUDA IPC / symmetric memory gives real ranks, minus NVLink.  Kernels
that wait on one another must never be co-scheduled on one GPU (B200_PROFILING.md), so the ranks' engines are advanced
in LOCKSTEP through ``LanczosEngine.step_phases``: all ranks launch phase k before any rank launches phase k+1, on one
stream, so every push has completed before the kernel that waits for it starts and no wait ever spins.
"""
import torch

from hessian_llm_vision_b200 import kernels, peer as peer_mod


class EmulatedComm:
    """world/rank without a process group: the peer path needs nothing else from it."""
    backend = "emulated"
    group = None

    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def barrier(self):
        pass

    def all_reduce_sum(self, t):
        raise AssertionError("the peer exchange must not fall back to collectives")

    reduce_scatter_sum = all_gather = all_reduce_sum


def make_contexts(world, n_pad, device):
    xchg = [torch.empty(kernels.peer_xchg_bytes(), dtype=torch.uint8, device=device) for _ in range(world)]
    hv = [torch.zeros(n_pad, dtype=torch.float32, device=device) for _ in range(world)]
    v = [torch.zeros(n_pad, dtype=torch.float32, device=device) for _ in range(world)]
    for x in xchg:
        kernels.peer_xchg_init(x)
    ptr = lambda ts: [t.data_ptr() for t in ts]
    return [peer_mod.make_context(world, r, xchg[r], hv[r], v[r], ptr(xchg), ptr(hv), ptr(v), spin_timeout_ms=2000)
            for r in range(world)]


def lockstep(engines, j):
    gens = [e.step_phases(j) for e in engines]
    live = True
    while live:
        live = False
        for g in gens:
            try:
                next(g)
                live = True
            except StopIteration:
                pass
