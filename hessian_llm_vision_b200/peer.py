"""Peer-memory plumbing for the sharded recurrence: who maps whose buffers.

The exchange itself happens inside libhlv kernels (csrc/hlv_peer.cuh); this module only produces a
``PeerContext`` -- for every rank, the addresses of three of its buffers AS MAPPED IN THIS PROCESS:

    xchg      the scalar / flag exchange area (hlv_peer_xchg_bytes() bytes)
    hv_full   the full-length Hessian-vector product each rank's HVP writes (peers read their shard of it)
    v_full    the full-length Lanczos vector each rank's HVP reads (peers write their shard into it)

``connect`` obtains them from torch's symmetric-memory allocator (CUDA VMM handles exchanged over the process
group's store; NVLink / NVSwitch peer mappings) -- plumbing, like the NCCL communicator it replaces.  Any other
mapping mechanism (CUDA IPC handles, a custom allocator) can fill a PeerContext the same way.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib, kernels


@dataclass
class PeerContext:
    world: int
    rank: int
    xchg: torch.Tensor                 # local exchange area (uint8)
    hv_full: torch.Tensor              # local full-length Hv (float32, n_pad)
    v_full: torch.Tensor               # local full-length v (float32, n_pad)
    ctx: _lib.PeerCtx                  # hlv_peer_ctx
    hv_ptrs: C.Array                   # void*[world]: every rank's hv_full
    v_ptrs: C.Array                    # void*[world]: every rank's v_full
    keepalive: Optional[list] = None   # whatever owns the mappings
    hv_multicast: int = 0              # NVSwitch multicast address of hv_full / v_full (0 = none): in-switch reduction
    v_multicast: int = 0               # and replicated stores (multimem.ld_reduce / multimem.st) instead of per-peer access

    def error(self) -> int:
        return kernels.peer_error(self)


def make_context(world: int, rank: int, xchg: torch.Tensor, hv_full: torch.Tensor, v_full: torch.Tensor,
                 xchg_ptrs: List[int], hv_ptrs: List[int], v_ptrs: List[int], spin_timeout_ms: int = 20000,
                 keepalive=None, hv_multicast: int = 0, v_multicast: int = 0) -> PeerContext:
    if not (len(xchg_ptrs) == len(hv_ptrs) == len(v_ptrs) == world) or world > _lib.HLV_MAX_PEERS:
        raise ValueError(f"need {world} pointers per buffer (max {_lib.HLV_MAX_PEERS} ranks)")
    ctx = _lib.PeerCtx()
    ctx.world, ctx.rank, ctx.spin_timeout_ms = world, rank, int(spin_timeout_ms)
    for p in range(world):
        ctx.xchg[p] = xchg_ptrs[p]
    return PeerContext(world=world, rank=rank, xchg=xchg, hv_full=hv_full, v_full=v_full, ctx=ctx,
                       hv_ptrs=(C.c_void_p * world)(*hv_ptrs), v_ptrs=(C.c_void_p * world)(*v_ptrs), keepalive=keepalive,
                       hv_multicast=int(hv_multicast), v_multicast=int(v_multicast))


def connect(comm, device, n_pad: int, spin_timeout_ms: int = 20000, multicast: bool = True) -> PeerContext:
    """Allocate the three buffers in symmetric memory and map every rank's copy (collective over ``comm``'s group).
    Raises if symmetric memory is not available; the engine then keeps its torch.distributed collectives."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm
    group = comm.group if comm.group is not None else dist.group.WORLD
    try:                                                    # older releases need the group announced first
        symm.enable_symm_mem_for_group(group.group_name)
    except Exception:  # noqa: BLE001
        pass
    dev = torch.device(device)
    bufs, handles = [], []
    for shape, dtype in (((kernels.peer_xchg_bytes(),), torch.uint8), ((n_pad,), torch.float32), ((n_pad,), torch.float32)):
        t = symm.empty(*shape, dtype=dtype, device=dev)
        h = symm.rendezvous(t, group)
        bufs.append(t)
        handles.append(h)
    xchg, hv_full, v_full = bufs
    ptrs = [[int(p) for p in h.buffer_ptrs] for h in handles]
    if any(len(p) != comm.world or any(x == 0 for x in p) for p in ptrs):
        raise RuntimeError("symmetric memory rendezvous did not return a mapping for every rank")
    kernels.peer_xchg_init(xchg)
    hv_full.zero_()
    v_full.zero_()
    torch.cuda.synchronize(dev)
    dist.barrier(group=group)                               # every area is zeroed before anybody pushes into it
    mc = [0, 0]
    if multicast:
        try:
            mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in handles[1:]]
        except Exception:  # noqa: BLE001
            mc = [0, 0]
    return make_context(comm.world, comm.rank, xchg, hv_full, v_full, ptrs[0], ptrs[1], ptrs[2], spin_timeout_ms,
                        keepalive=handles, hv_multicast=mc[0], v_multicast=mc[1])
