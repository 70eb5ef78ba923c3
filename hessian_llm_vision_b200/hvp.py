"""Hessian-vector-product operators over a model's flattened parameters.

The HVP itself stays a torch double-backward (BASELINE north star); what this module adds
around it is the flat-vector plumbing done by libhlv kernels instead of 148 slices,
444 tiny dot kernels and a torch.cat per application:

  * scatter: ``v`` is sliced into zero-copy per-parameter views (no kernel needed while the
    flat vector is fp32 and on the parameter device);
  * second backward as ``autograd.grad(grads, params, grad_outputs=v_views)`` -- the same
    derivative as the reference's ``sum(v*g).backward()`` (gpt2_hessian_cpu.py:103-106)
    without materialising the dot;
  * gather: the per-parameter results go straight into the flat Lanczos vector with
    ``hlv_gather_f32`` (bit-exact torch.cat replacement, gpt2_hessian_cpu.py:109), accumulated
    across micro-batches, with alpha = <Hv, v> fused into the last pass.

Operator flavours of the reference that are covered:
  full model, one batch ............ hess_vec,            gpt2_hessian_cpu.py:75-109
  full model, dataset average ...... hess_vec,            gpt2_savehessian.py:130-163
  one transformer block ............ hess_vec_layer,      ipynbs/visual-eigen.ipynb cell 10
  block-diagonal per tensor ........ hess_vec_layer_by_layer, gpt2_savehessian_layer.py:130-178
  CIFAR nets (criterion, BN train) . hess_vec,            train_savespec.py:61-91
"""
from __future__ import annotations

import contextlib
from typing import Callable, Iterable, List, Optional, Sequence

import torch


def lm_loss(model, batch):
    """``model(input_ids=ids, labels=ids).loss`` -- labels are the inputs, no padding mask
    (gpt2_hessian_cpu.py:94-97; quirk Q9)."""
    ids = batch["input_ids"] if isinstance(batch, dict) else batch
    loss = model(input_ids=ids, labels=ids, use_cache=False).loss      # no KV cache: same loss, nothing to allocate
    return loss.mean() if loss.dim() > 0 else loss


def criterion_loss(criterion: Callable):
    """``criterion(model(x), y)`` for (input, target) batches (train_savespec.py:80-81)."""
    def _loss(model, batch):
        x, y = batch
        return criterion(model(x), y)
    return _loss


def _batch_size(batch) -> int:
    if isinstance(batch, dict):
        return int(batch["input_ids"].shape[0])
    if isinstance(batch, (tuple, list)):
        return int(batch[0].shape[0])
    return int(batch.shape[0])


def _to_device(batch, device, non_blocking=True):
    if isinstance(batch, dict):
        return {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in batch.items()}
    if isinstance(batch, (tuple, list)):
        return type(batch)(b.to(device, non_blocking=non_blocking) if torch.is_tensor(b) else b for b in batch)
    return batch.to(device, non_blocking=non_blocking)


def _math_sdpa():
    """Double-backward needs the MATH scaled-dot-product backend (flash / mem-efficient
    backward have no derivative)."""
    try:
        from torch.nn.attention import SDPBackend, sdpa_kernel
        return sdpa_kernel(SDPBackend.MATH)
    except Exception:  # pragma: no cover
        return contextlib.nullcontext()


@contextlib.contextmanager
def _capturable_scalar_tensors():
    """``torch.tensor(<python scalar>, device=cuda)`` is a pageable host-to-device copy, which is illegal while
    a CUDA graph is being captured -- and transformers' eager-attention mask builder does exactly that
    (masking_utils.eager_mask: ``torch.tensor(0.0, device=mask.device, dtype=dtype)``).  While capturing, such
    0-dim constants are produced by a fill kernel instead (same value, same dtype rules)."""
    orig = torch.tensor

    def tensor(data, *args, **kwargs):
        dev = kwargs.get("device")
        if isinstance(data, (bool, int, float)) and not args and dev is not None and torch.device(dev).type == "cuda":
            dtype = kwargs.get("dtype")
            if dtype is None:
                dtype = torch.bool if isinstance(data, bool) else (torch.int64 if isinstance(data, int) else torch.get_default_dtype())
            out = torch.full((), data, dtype=dtype, device=dev)
            return out.requires_grad_(True) if kwargs.get("requires_grad") else out
        return orig(data, *args, **kwargs)

    torch.tensor = tensor
    try:
        yield
    finally:
        torch.tensor = orig


def shard_batches(batches: Sequence, rank: int, world: int) -> List:
    """Round-robin micro-batch assignment for the batch-sharded HVP (one process per GPU)."""
    return [b for i, b in enumerate(batches) if i % world == rank]


class HessianVectorProduct:
    """H v for the mean loss over ``batches`` w.r.t. ``params`` (default: all of
    ``model.parameters()``, the reference's flat order).

    weights[i] multiplies micro-batch i's loss; the default B_i / N reproduces the dataset mean
    (diego_pythia.py:114, train_savespec.py:82).  With batch sharding pass ``total_sequences`` =
    the GLOBAL number of sequences so that the sum over ranks is the global mean.

    per_tensor=True gives the block-diagonal-by-tensor operator of gpt2_savehessian_layer.py.

    cache_graph: the forward pass and the first backward (gpt2_hessian_cpu.py:94-102) do not depend on v, so
    their autograd graph can be built ONCE per operator and every application then costs only the second
    backward (the reference rebuilds it every time).  The result is bit-identical (same kernels on the same
    saved tensors; tests/test_gpu_zz_full_size.py).  ``"auto"`` (default) keeps a micro-batch's graph when the
    device has room for it and for the second backward's temporaries, ``True`` always, ``False`` never.  A
    cached graph is valid while the model's weights and the batches do not change: build a new operator per
    spectrum run (as the reference scripts do) or call ``clear_cache()``.
    """

    def __init__(self, model: torch.nn.Module, batches: Iterable, loss_fn: Callable = lm_loss,
                 params: Optional[Sequence[torch.nn.Parameter]] = None,
                 weights: Optional[Sequence[float]] = None, total_sequences: Optional[int] = None,
                 per_tensor: bool = False, bn_train_mode: bool = False, cache_graph="auto",
                 device: Optional[torch.device] = None):
        self.model = model
        self.params = list(model.parameters()) if params is None else list(params)
        if not self.params:
            raise ValueError("no parameters to differentiate")
        self.device = torch.device(device) if device is not None else self.params[0].device
        self.numels = [p.numel() for p in self.params]
        self.n = sum(self.numels)
        self.batches = list(batches)
        if not self.batches:
            raise ValueError("need at least one batch")
        if weights is None:
            sizes = [_batch_size(b) for b in self.batches]
            tot = float(total_sequences if total_sequences is not None else sum(sizes))
            weights = [s / tot for s in sizes]
        self.weights = [float(w) for w in weights]
        self.loss_fn = loss_fn
        self.per_tensor = per_tensor
        self.bn_train_mode = bn_train_mode
        if cache_graph not in (True, False, "auto"):
            raise ValueError("cache_graph must be True, False or 'auto'")
        self.cache_graph = cache_graph
        self._graphs = {}
        self.cached_bytes = 0                       # device memory held by the kept first-backward graphs
        self.first_backward_builds = 0
        self.applications = 0
        self.h2d_bytes = 0

    # -- pieces -------------------------------------------------------------------
    def _views(self, v: torch.Tensor) -> List[torch.Tensor]:
        v = v.detach().reshape(-1)
        if v.numel() != self.n:
            raise ValueError(f"vector has {v.numel()} elements, operator dimension is {self.n}")
        if v.device != self.device:
            v = v.to(self.device)                   # reference: .to(param.device), gpt2_hessian_cpu.py:81
        return [s.view_as(p) for s, p in zip(torch.split(v, self.numels), self.params)]

    def _prepare_model(self):
        self.model.eval()                           # gpt2_hessian_cpu.py:84
        if self.bn_train_mode:                      # :85-86
            for m in self.model.modules():
                if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                    m.train()

    def _first_backward(self, i: int):
        if self.cache_graph and i in self._graphs:
            return self._graphs[i]
        batch = self.batches[i]
        probe = batch["input_ids"] if isinstance(batch, dict) else (batch[0] if isinstance(batch, (tuple, list)) else batch)
        if probe.device != self.device:
            self.h2d_bytes += sum(t.numel() * t.element_size() for t in
                                  (batch.values() if isinstance(batch, dict) else (batch if isinstance(batch, (tuple, list)) else [batch]))
                                  if torch.is_tensor(t))
            batch = _to_device(batch, self.device)
        on_cuda = self.device.type == "cuda"
        before = torch.cuda.memory_allocated(self.device) if on_cuda and self.cache_graph else 0
        loss = self.loss_fn(self.model, batch) * self.weights[i]
        grads = torch.autograd.grad(loss, self.params, create_graph=True, allow_unused=True)
        self.first_backward_builds += 1
        if self.cache_graph == "auto":              # decided once, on the first graph built
            self.cache_graph = self._room_to_keep(torch.cuda.memory_allocated(self.device) - before) if on_cuda else True
        if self.cache_graph:
            self._graphs[i] = grads
            if on_cuda:
                self.cached_bytes += torch.cuda.memory_allocated(self.device) - before
        return grads

    def _room_to_keep(self, graph_bytes: int) -> bool:
        """Keep the first-backward graphs only if every micro-batch's graph plus the second backward's temporaries
        (about as large as one graph again) fit in what the device has left."""
        free, _ = torch.cuda.mem_get_info(self.device)
        free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        remaining = len(self.batches) - 1
        return free >= (remaining + 1.5) * max(graph_bytes, 0)

    def pieces(self, v: torch.Tensor, i: int) -> List[torch.Tensor]:
        """Per-parameter pieces of (weights[i] * H_i) v for micro-batch i."""
        views = self._views(v)
        self._prepare_model()
        with _math_sdpa():
            grads = self._first_backward(i)
            keep = i in self._graphs
            if not self.per_tensor:
                used = [(g, p, x) for g, p, x in zip(grads, self.params, views) if g is not None and g.requires_grad]
                hv = torch.autograd.grad([g for g, _, _ in used], [p for _, p, _ in used],
                                         grad_outputs=[x for _, _, x in used], retain_graph=keep, allow_unused=True)
                by_param = {id(p): h for (_, p, _), h in zip(used, hv)}
                out = [by_param.get(id(p)) for p in self.params]
            else:
                out = []
                last = max((k for k, g in enumerate(grads) if g is not None and g.requires_grad), default=-1)
                for k, (g, p, x) in enumerate(zip(grads, self.params, views)):
                    if g is None or not g.requires_grad:
                        out.append(None)
                        continue
                    h = torch.autograd.grad(g, p, grad_outputs=x, retain_graph=keep or k != last, allow_unused=True)[0]
                    out.append(h)
        return [h.contiguous() if h is not None else torch.zeros_like(p) for h, p in zip(out, self.params)]

    # -- flat results -----------------------------------------------------------------
    def accumulate_into(self, v: torch.Tensor, out: torch.Tensor, dot_with=None, dot_out=None,
                        ws=None, ops=None, phases=None) -> None:
        """out[:n] = H v (sum over this operator's micro-batches), gathered by libhlv;
        optionally dot_out = <out, dot_with> fused into the last gather."""
        if ops is None:
            from . import kernels as ops
        nb = len(self.batches)
        for i in range(nb):
            pcs = self.pieces(v, i)
            last = i == nb - 1
            if phases is not None:
                phases.start("gather")
            ops.gather(pcs, out, accumulate=i > 0, dot_with=dot_with if last else None,
                       dot_out=dot_out if last else None, ws=ws)
            if phases is not None:
                phases.stop("gather")
        self.applications += 1

    def __call__(self, v: torch.Tensor) -> torch.Tensor:
        """Drop-in ``hess_vec``: flat [P] in -> flat [P] out ([P,1] -> [P,1])."""
        col = v.dim() == 2
        out = torch.empty(self.n, dtype=torch.float32, device=self.device)
        self.accumulate_into(v, out)
        return out.unsqueeze(1) if col else out

    def clear_cache(self) -> None:
        """Drop the kept first-backward graphs (after the model's weights or the batches changed)."""
        self._graphs.clear()
        self.cached_bytes = 0

    def capture(self, ws=None, warmup: int = 3, out: Optional[torch.Tensor] = None, pipeline: bool = False,
                reuse_first: Optional[bool] = None) -> "GraphedHVP":
        """Capture the application into CUDA graphs: none of the ~4,000 kernel launches per double-backward is
        issued from Python any more (worth most when the per-rank batch is small and the launches are host-bound).
        Batches may live in pinned host memory: their H2D copies become nodes of the graph.

        Two graphs over one memory pool: the v-independent half (every micro-batch's forward + first backward) and
        the v-dependent half (second backward + libhlv gather + fused <Hv, v>).
        ``reuse_first`` (default: the operator's ``cache_graph`` setting): the first half is replayed ONCE -- at the
        first application and again after ``invalidate()`` -- and every application replays only the second half.
        ``reuse_first=False`` redoes ALL of the work on every application, like the reference; with
        ``pipeline=True`` the first half of the NEXT application is then started on a side stream as soon as this
        one has finished, so it overlaps whatever the caller does in between (the Lanczos recurrence and its
        collectives).  The operator (weights, batches) must not change while a graph's first half is live -- call
        ``invalidate()`` if it does.
        ``out``: write H v straight into this buffer (e.g. the engine's w) instead of a private one."""
        return GraphedHVP(self, ws=ws, warmup=warmup, out=out, pipeline=pipeline, reuse_first=reuse_first)


class GraphedHVP:
    """A HessianVectorProduct application replayed from CUDA graphs.

    Static buffers: ``v`` (input, length n), ``out`` (H v, length n) and ``dot`` (<Hv, v>).  The
    engine passes its own vector; it is copied into the static input (one 4n-byte D2D copy), the
    graph is replayed, and the result is copied out (or lands in the caller's buffer given as ``out``)."""

    def __init__(self, op: "HessianVectorProduct", ws=None, warmup: int = 3, out: Optional[torch.Tensor] = None,
                 pipeline: bool = False, reuse_first: Optional[bool] = None):
        from . import kernels
        for b in op.batches:
            t = _first(b)
            if t.device != op.device and not (t.device.type == "cpu" and t.is_pinned()):
                raise ValueError("capture() needs device-resident or pinned-host batches")
        self.op = op
        self.n = op.n
        dev = op.device
        self.ws = ws if ws is not None else kernels.Workspace(dev, max_rows=1)
        self.v = torch.zeros(op.n, dtype=torch.float32, device=dev)
        if out is not None:
            if out.device != dev or out.dtype != torch.float32 or out.numel() < op.n or not out.is_contiguous():
                raise ValueError("capture(out=...): need a contiguous float32 device buffer of at least n elements")
            self.out = out.reshape(-1)[: op.n]
        else:
            self.out = torch.zeros(op.n, dtype=torch.float32, device=dev)
        self.dot = torch.zeros(1, dtype=torch.float64, device=dev)
        self.v.normal_()
        self.v /= torch.linalg.vector_norm(self.v)

        if reuse_first is None:
            reuse_first = op.cache_graph is not False and not pipeline
        if reuse_first and pipeline:
            raise ValueError("capture(): pipeline prefetches a first half that reuse_first would not replay; pick one")
        self.reuse_first = bool(reuse_first)
        self.pipeline = bool(pipeline)
        self.split = self.pipeline or self.reuse_first      # two graphs (first half / second half) or one
        self.graph_first = None
        self._primed = False
        self.first_replays = 0

        def run():
            op.accumulate_into(self.v, self.out, dot_with=self.v, dot_out=self.dot, ws=self.ws, ops=kernels)

        def first_half():                       # forward + first backward of every micro-batch: does not depend on v
            op._prepare_model()
            with _math_sdpa():
                for i in range(len(op.batches)):
                    op._first_backward(i)
        # the capture decides the caching policy itself (no memory query while a capture is open)
        op.cache_graph = True if self.split else False   # split: run() performs only the second backward over first_half()'s graphs
        # A cached first-backward graph must be (re)built on the side stream: autograd replays each node on
        # the stream its forward ran on, and a capture may fork into a side stream but never into the legacy
        # default stream (cudaErrorStreamCaptureInvalidated).
        op.clear_cache()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), _capturable_scalar_tensors():
            for _ in range(max(1, warmup)):
                if self.split:
                    op.clear_cache()
                    first_half()
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        before, h2d0 = kernels.launch_count, op.h2d_bytes
        if self.split:
            # two graphs over ONE memory pool: the second backward reads the activations / first-order gradients the
            # first graph leaves behind; they are always replayed in capture order and never concurrently
            op.clear_cache()
            pool = torch.cuda.graph_pool_handle()
            self.graph_first = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_first, pool=pool), _capturable_scalar_tensors():
                first_half()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, pool=pool), _capturable_scalar_tensors():
                run()
            self.side = torch.cuda.Stream(device=dev)
            self._ev_first = torch.cuda.Event()
            self._ev_second = torch.cuda.Event()
        else:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph), _capturable_scalar_tensors():
                run()
        self.launches_per_replay = kernels.launch_count - before
        self.h2d_bytes_per_replay = op.h2d_bytes - h2d0     # pinned-host batches: copied by the graph that runs the forward pass
        op.h2d_bytes = h2d0                                  # capture and warm-up are setup, not applications
        self.applications = 0

    # -- the v-independent half ------------------------------------------------------------------------------------
    def _launch_first(self, main) -> None:
        """pipelined mode: next application's forward + first backward on the side stream"""
        self._ev_second.record(main)                        # everything this application read is finished with
        self.side.wait_event(self._ev_second)
        with torch.cuda.stream(self.side):
            self.graph_first.replay()
            self._ev_first.record(self.side)
        self.op.h2d_bytes += self.h2d_bytes_per_replay
        self.first_replays += 1
        self._primed = True

    def drain(self) -> None:
        """Make the current stream wait for a prefetched first half (so a timed region accounts for it)."""
        if self.pipeline and self._primed:
            torch.cuda.current_stream(self.op.device).wait_event(self._ev_first)

    def invalidate(self) -> None:
        """The first half must be redone before the next application (call after the model's weights or the
        batches changed, or to account a run's first-backward build inside a timed region)."""
        if self.pipeline and self._primed:
            torch.cuda.current_stream(self.op.device).wait_event(self._ev_first)
        self._primed = False

    @property
    def weights(self):
        return self.op.weights

    @property
    def h2d_bytes(self):
        return self.op.h2d_bytes

    def accumulate_into(self, v: torch.Tensor, out: torch.Tensor, dot_with=None, dot_out=None,
                        ws=None, ops=None, phases=None) -> None:
        from . import kernels
        v = v.reshape(-1)
        if v.numel() != self.n or v.device != self.op.device:
            raise ValueError(f"captured operator takes a length-{self.n} vector on {self.op.device}, got {v.numel()} on {v.device}")
        main = torch.cuda.current_stream(self.op.device)
        if self.pipeline:
            if not self._primed:
                self._launch_first(main)
            main.wait_event(self._ev_first)
        elif self.reuse_first:
            if not self._primed:                # first application (or after invalidate()): forward + first backward, once
                self.graph_first.replay()
                self.op.h2d_bytes += self.h2d_bytes_per_replay
                self.first_replays += 1
                self._primed = True
        else:
            self.op.h2d_bytes += self.h2d_bytes_per_replay
        self.v.copy_(v)
        self.graph.replay()
        kernels.launch_count += self.launches_per_replay
        if out.data_ptr() != self.out.data_ptr():
            out.copy_(self.out)
        if dot_out is not None:
            if dot_with is None or dot_with.data_ptr() == v.data_ptr():
                dot_out.copy_(self.dot)       # the captured <Hv, v> (alpha in the Lanczos loop)
            else:                             # a dot with some other vector is not what the graph computed
                kernels.dot(out.reshape(-1)[: self.n], dot_with.reshape(-1), dot_out, ws if ws is not None else self.ws)
        if self.pipeline:
            self._launch_first(main)          # next application's forward + first backward, overlapping the caller's work
        self.applications += 1

    def __call__(self, v: torch.Tensor) -> torch.Tensor:
        col = v.dim() == 2
        out = torch.empty(self.n, dtype=torch.float32, device=self.op.device)
        self.accumulate_into(v, out)
        return out.unsqueeze(1) if col else out


def _first(batch):
    if isinstance(batch, dict):
        return batch["input_ids"]
    if isinstance(batch, (tuple, list)):
        return batch[0]
    return batch


class CurvVecProduct:
    """Same constructor and call shape as the reference's adapter class
    (gpt2_savehessian.py:166-189): ``CurvVecProduct(loader, model, init_vec=None)``, called with
    a [P,1] column and returning a [P,1] column ON THE DEVICE (no ``.cpu()``, cf.
    gpt2_hessian_cpu.py:137 vs gpt2_hessian_gpu.py:137).

    The reference swaps ``init_vec`` in on the first call while gpytorch keeps its own random
    q_0 (quirk F3).  Here ``init_vec`` is exposed as an attribute and ``lanczos_tridiag`` uses
    it as the actual start vector, which is what the scripts intend."""

    def __init__(self, loader, model, init_vec: Optional[torch.Tensor] = None, criterion=None,
                 layer: Optional[torch.nn.Module] = None, **kwargs):
        loss_fn = criterion_loss(criterion) if criterion is not None else lm_loss
        params = list(layer.parameters()) if layer is not None else None
        self.op = HessianVectorProduct(model, loader, loss_fn=loss_fn, params=params, **kwargs)
        self.init_vec = init_vec
        self.iters = 0

    @property
    def n(self) -> int:
        return self.op.n

    def accumulate_into(self, *a, **k):
        self.iters += 1
        return self.op.accumulate_into(*a, **k)

    def __call__(self, vector: torch.Tensor) -> torch.Tensor:
        self.iters += 1
        out = self.op(vector.reshape(-1))
        return out.unsqueeze(1)
