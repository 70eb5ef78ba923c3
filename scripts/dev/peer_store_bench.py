"""Dev probe (torchrun, N ranks): the peer stores of hlv_x_normalize_store_f32 and the peer loads of
hlv_x_reduce_scatter_dot_f32 alone, unicast vs multicast, for a few grid sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import hessian_llm_vision_b200 as hlv
from hessian_llm_vision_b200 import kernels as K, peer as P

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
comm = hlv.Comm()
n = 124_046_592
sn = (-(-n // world) + 7) // 8 * 8
ctx = P.connect(comm, dev, sn * world)
ws = K.Workspace(dev, max_rows=4)
w = torch.randn(sn, device=dev)
vout = torch.empty(sn, device=dev)
nrm = torch.ones(1, dtype=torch.float64, device=dev)
beta = torch.zeros(1, dtype=torch.float64, device=dev)
alpha = torch.zeros(1, dtype=torch.float64, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def norm_once(mc):
    # feed HLV_CH_NORM so the pull finds its epoch, then normalise + store into every rank
    K.x_cgs_update(ctx, torch.zeros(1, sn, device=dev), 1, torch.zeros(1, dtype=torch.float64, device=dev), wz, nrm_scratch, ws)


wz = torch.zeros(sn, device=dev)
for grid in (0,):   # the grid is fixed in the library now (one CTA per SM for peer stores)
    os.environ["HLV_DEV_NORM_GRID"] = str(grid)
    for mc in (0, 1):
        def f():
            K.x_normalize_store(ctx, w, nrm, beta, vout, None, ctx.v_ptrs, rank * sn, 0.0, None, 0, ws,
                                v_multicast=ctx.v_multicast if mc else 0)
            K.peer_wait(ctx, 5)
        t = timeit(f)
        if rank == 0:
            print(f"normalize+stores world={world} grid={grid or 'auto'} multicast={mc}: {t:.3f} ms  remote {4 * sn * (1 if mc else world - 1) / t / 1e6:.0f} GB/s egress", flush=True)
for rsg in (0,):
  for mc in (0, 1):
    def g():
        K.peer_signal(ctx, 0)
        K.x_reduce_scatter_dot(ctx, ctx.hv_ptrs, rank * sn, wz, w, alpha, ws, hv_multicast=ctx.hv_multicast if mc else 0)
    t = timeit(g)
    if rank == 0:
        print(f"reduce_scatter+alpha world={world} grid={rsg or 'auto'} multicast={mc}: {t:.3f} ms  remote {4 * sn * (world - 1) / t / 1e6:.0f} GB/s ingress-equivalent", flush=True)
# NCCL reference points
full = torch.empty(sn * world, device=dev)
t = timeit(lambda: dist.all_gather_into_tensor(full, vout))
if rank == 0:
    print(f"nccl all_gather: {t:.3f} ms", flush=True)
t = timeit(lambda: dist.reduce_scatter_tensor(wz, full))
if rank == 0:
    print(f"nccl reduce_scatter: {t:.3f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
