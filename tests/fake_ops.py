"""TEST DOUBLE ONLY.  An oracle-backed (torch CPU) stand-in with the same call surface as
``hessian_llm_vision_b200.kernels``.  It exists so the HOST logic of the engine -- iteration
order, buffer rotation, sharding along the parameter dimension, the k-float all-reduces --
can be exercised on CPU under gloo with world_size 2.  The product never uses it: the engine
defaults to the CUDA library and refuses CPU tensors."""
import torch

launch_count = 0


def install(monkeypatch=None):
    """Point the engine at this double.  With a pytest ``monkeypatch`` the swap is undone after the test; a
    worker subprocess (gloo tests) passes nothing and keeps it for its lifetime."""
    import sys
    import hessian_llm_vision_b200  # noqa: F401  (the package exports a FUNCTION named lanczos; fetch the module itself)
    lz = sys.modules["hessian_llm_vision_b200.lanczos"]
    me = sys.modules[__name__]
    if monkeypatch is not None:
        monkeypatch.setattr(lz, "_kernels", me)
    else:
        lz._kernels = me


def require_device(device):
    pass


def compute_device(requested):
    return torch.device(requested)


class Workspace:
    def __init__(self, device, max_rows=128):
        self.max_rows = max_rows
        self.nbytes = 0
        self.ptr = 0


class TensorList(list):
    pass


def gather(tensors, dst, *, scale=1.0, accumulate=False, dot_with=None, dot_out=None, ws=None):
    flat = torch.cat([t.reshape(-1) for t in tensors])
    if scale == 1.0 and not accumulate:
        dst.copy_(flat)
    else:
        new = (dst if accumulate else torch.zeros_like(dst)) + scale * flat
        dst.copy_(new)
    if dot_out is not None:
        dot_out[0] = torch.dot(dst.double(), dot_with.double())


def scatter(src, tensors):
    off = 0
    for t in tensors:
        t.copy_(src[off: off + t.numel()].view_as(t))
        off += t.numel()


def dot(a, b, out, ws):
    out[0] = torch.dot(a.double(), b.double())


def lanczos_update(w, vj, vjm1, alpha, beta, norm2_out, ws):
    a = alpha[0].float()
    if vjm1 is None:
        w -= a * vj
    else:
        w -= (a * vj + beta[0].float() * vjm1)
    norm2_out[0] = torch.dot(w.double(), w.double())


def normalize_store(w, norm2, beta_out, v_out, row_bf16=None, breakdown_tol=0.0, breakdown_iter=None, it=0):
    beta = torch.sqrt(norm2[0])
    beta_out[0] = beta
    if breakdown_iter is not None and float(beta) < breakdown_tol and int(breakdown_iter[0]) < 0:
        breakdown_iter[0] = it
    if v_out is not None or row_bf16 is not None:
        v = w / beta.float()
        if v_out is not None:
            v_out.copy_(v)
        if row_bf16 is not None:
            row_bf16.copy_(v.to(torch.bfloat16))


def cgs_project(V, rows, w, c_out, ws):
    c_out[:rows] = V[:rows, : w.numel()].double() @ w.double()


def cgs_needs_pass(c, rows, norm2, tol, flag_out):
    flag_out[0] = int(bool((c[:rows].abs() > tol * torch.sqrt(norm2[0])).any()))


def cgs_update(V, rows, c, w, norm2_out, ws, sign=-1.0, run_flag=None):
    if run_flag is not None and int(run_flag[0]) == 0:
        return
    w += sign * (V[:rows, : w.numel()].float().t() @ c[:rows].float())
    if norm2_out is not None:
        norm2_out[0] = torch.dot(w.double(), w.double())


def fused_max_rows(dtype):
    return 200 if dtype == torch.float32 else 400


def cgs_update_project(V, rows, c_in, w, c_out, norm2_out, ws):
    cgs_update(V, rows, c_in, w, norm2_out, ws)
    cgs_project(V, rows, w, c_out, ws)


def ritz_vectors(Q, m, Y, out, n):
    out[:, :n] = Y[:m].t() @ Q[:m, :n].float()


def x_update_project(peer, V, rows, w, vj, vjm1, alpha, beta, c_out, ws):
    assert peer is None
    a = alpha[0].float()
    if vjm1 is None:
        w -= a * vj
    else:
        w -= (a * vj + beta[0].float() * vjm1)
    cgs_project(V, rows, w, c_out, ws)
