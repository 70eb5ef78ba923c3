"""-m gpu: every libhlv kernel, called through the C ABI, against the CPU oracle."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

GPT2_SHAPES = ([(50257, 768), (512, 768)] +
               [s for _ in range(12) for s in [(768,), (768,), (768, 2304), (2304,), (768, 768), (768,), (768,), (768,),
                                               (768, 3072), (3072,), (3072, 768), (768,)]] +
               [(768,), (768,)])


@pytest.fixture(scope="module")
def K(cuda_dev, libhlv):
    from hessian_llm_vision_b200 import kernels
    return kernels


def _ws(K, dev, rows=128):
    return K.Workspace(dev, max_rows=rows)


# ------------------------------------------------------------------ (a) gather / scatter
def test_gather_scatter_gpt2_full_size_bit_exact(K, cuda_dev):
    """The real thing: GPT-2's 148 tensors, P = 124,046,592, vs torch.cat, bit for bit."""
    assert len(GPT2_SHAPES) == 148
    g = torch.Generator(device=cuda_dev).manual_seed(0)
    tensors = [torch.randn(*s, device=cuda_dev, generator=g) for s in GPT2_SHAPES]
    P = sum(t.numel() for t in tensors)
    assert P == 124_046_592
    v = torch.randn(P, device=cuda_dev, generator=g)
    dst = torch.empty(P, device=cuda_dev)
    dot = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    ws = _ws(K, cuda_dev)
    K.gather(tensors, dst, dot_with=v, dot_out=dot, ws=ws)
    ref = torch.cat([t.view(-1) for t in tensors])
    assert torch.equal(dst, ref)
    exact = torch.dot(ref.double(), v.double()).item()
    assert abs(dot.item() - exact) <= 1e-6 * float(torch.linalg.vector_norm(ref.double()) * torch.linalg.vector_norm(v.double()))
    # plain gather (no dot) and scatter back
    dst2 = torch.zeros(P, device=cuda_dev)
    K.gather(tensors, dst2)
    assert torch.equal(dst2, ref)
    outs = [torch.zeros_like(t) for t in tensors]
    K.scatter(dst2, outs)
    for a, b in zip(outs, tensors):
        assert torch.equal(a, b)


@pytest.mark.parametrize("sizes", [
    [1], [3, 5, 7], [0, 4, 0, 9, 1], [1023, 1, 4097, 2, 8191, 8193], [64, 3, 9408, 64, 64, 1000, 10],
    [17] * 300,                       # > 224 tensors: large parameter table
    [5] * 1500,                       # > 1024 tensors: chunked launches
])
def test_gather_ragged_unaligned(K, cuda_dev, sizes):
    """Odd sizes put every segment boundary off the 16-byte grid (ResNet-like); empty tensors; long lists."""
    g = torch.Generator(device=cuda_dev).manual_seed(1)
    tensors = [torch.randn(s, device=cuda_dev, generator=g) for s in sizes]
    P = sum(sizes)
    ref = torch.cat([t.view(-1) for t in tensors]) if P else torch.zeros(0, device=cuda_dev)
    v = torch.randn(P, device=cuda_dev, generator=g)
    dst = torch.full((P,), float("nan"), device=cuda_dev)
    dot = torch.full((1,), float("nan"), dtype=torch.float64, device=cuda_dev)
    ws = _ws(K, cuda_dev)
    K.gather(tensors, dst, dot_with=v, dot_out=dot, ws=ws)
    assert torch.equal(dst, ref)
    assert abs(dot.item() - torch.dot(ref.double(), v.double()).item()) < 1e-4 * (1 + P ** 0.5)
    outs = [torch.full_like(t, float("nan")) for t in tensors]
    K.scatter(dst, outs)
    for a, b in zip(outs, tensors):
        assert torch.equal(a, b)


def test_gather_views_with_offset_sources(K, cuda_dev):
    """Sources that are 4-byte but not 16-byte aligned (views into a larger buffer)."""
    big = torch.randn(10_000, device=cuda_dev)
    tensors = [big[1:1001], big[2003:2010], big[3002:7003]]
    dst = torch.empty(sum(t.numel() for t in tensors), device=cuda_dev)
    K.gather([t.contiguous() if not t.is_contiguous() else t for t in tensors], dst)
    assert torch.equal(dst, torch.cat(tensors))


def test_gather_scale_accumulate(K, cuda_dev):
    g = torch.Generator(device=cuda_dev).manual_seed(2)
    sizes = [1000, 33, 4096, 7]
    a = [torch.randn(s, device=cuda_dev, generator=g) for s in sizes]
    b = [torch.randn(s, device=cuda_dev, generator=g) for s in sizes]
    P = sum(sizes)
    v = torch.randn(P, device=cuda_dev, generator=g)
    dst = torch.empty(P, device=cuda_dev)
    dot = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    ws = _ws(K, cuda_dev)
    K.gather(a, dst)
    K.gather(b, dst, scale=0.25, accumulate=True, dot_with=v, dot_out=dot, ws=ws)
    ref = torch.cat(a) + 0.25 * torch.cat(b)
    assert torch.allclose(dst, ref, rtol=1e-6, atol=1e-7)
    assert abs(dot.item() - torch.dot(dst.double(), v.double()).item()) < 1e-3


# ------------------------------------------------------------------ (b) recurrence
@pytest.mark.parametrize("n", [1, 7, 1024, 4099, 1_000_003, 8 * 1024 * 1024])
def test_dot_update_normalize(K, cuda_dev, n):
    g = torch.Generator(device=cuda_dev).manual_seed(n)
    w = torch.randn(n, device=cuda_dev, generator=g)
    vj = torch.randn(n, device=cuda_dev, generator=g)
    vo = torch.randn(n, device=cuda_dev, generator=g)
    ws = _ws(K, cuda_dev)
    out = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    K.dot(w, vj, out, ws)
    exact = torch.dot(w.double(), vj.double()).item()
    assert abs(out.item() - exact) <= 2e-6 * (n ** 0.5) * 3 + 1e-6
    # three-term update: bit-identical to torch's elementwise sequence given the same scalars
    alpha = torch.tensor([0.7310585], dtype=torch.float64, device=cuda_dev)
    beta = torch.tensor([1.6180339], dtype=torch.float64, device=cuda_dev)
    nrm = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    w1 = w.clone()
    K.lanczos_update(w1, vj, vo, alpha, beta, nrm, ws)
    ref = w - (alpha.float() * vj + beta.float() * vo)          # lanczostrain_hand.py:202
    assert torch.equal(w1, ref)
    assert abs(nrm.item() - torch.dot(ref.double(), ref.double()).item()) <= 1e-6 * nrm.item() + 1e-9
    w0 = w.clone()
    K.lanczos_update(w0, vj, None, alpha, None, nrm, ws)
    assert torch.equal(w0, w - alpha.float() * vj)               # :185
    # normalise + store: beta = sqrt(norm2), v = w / beta with true division
    K.lanczos_update(w1, vj, vo, torch.zeros_like(alpha), torch.zeros_like(beta), nrm, ws)   # nrm = |w1|^2
    beta_out = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    n8 = (n + 7) // 8 * 8
    v_out = torch.zeros(n8, device=cuda_dev)[:n]
    row16 = torch.zeros(n8, dtype=torch.bfloat16, device=cuda_dev)[:n]
    flag = torch.full((1,), -1, dtype=torch.int32, device=cuda_dev)
    K.normalize_store(w1, nrm, beta_out, v_out, row16, 0.0, flag, 3)
    b32 = beta_out.float()
    assert abs(beta_out.item() - torch.linalg.vector_norm(w1.double()).item()) <= 1e-6 * beta_out.item()
    assert torch.equal(v_out, w1 / b32)                          # :193
    assert torch.equal(row16, (w1 / b32).to(torch.bfloat16))
    assert flag.item() == -1
    K.normalize_store(w1, nrm, beta_out, None, None, 1e30, flag, 5)    # beta-only + breakdown flag
    assert flag.item() == 5


# ------------------------------------------------------------------ (c) CGS project / update
def _basis(rows, n, dev, dtype, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    ld = (n + 7) // 8 * 8
    V = torch.zeros(rows, ld, device=dev, dtype=dtype)
    V[:, :n] = (torch.randn(rows, n, device=dev, generator=g) / n ** 0.5).to(dtype)
    w = torch.randn(n, device=dev, generator=g)
    return V, w


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,n", [(1, 1), (1, 2048), (3, 5), (8, 2048), (9, 4096), (13, 2049), (100, 65536 + 17),
                                    (37, 1_000_003), (128, 300_000), (5, 23_528_522)])
def test_cgs_project_update(K, cuda_dev, dtype, rows, n):
    V, w = _basis(rows, n, cuda_dev, dtype, rows * 7 + n)
    ws = _ws(K, cuda_dev, rows)
    c = torch.full((rows,), float("nan"), dtype=torch.float64, device=cuda_dev)
    K.cgs_project(V, rows, w, c, ws)
    Vd = V[:, :n].double()
    c_ref = Vd @ w.double()
    tol = 3e-6 * float((Vd.abs() @ w.double().abs()).max()) + 1e-12
    assert float((c - c_ref).abs().max()) <= tol
    # update (sign=-1) + norm
    w1 = w.clone()
    nrm = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    K.cgs_update(V, rows, c, w1, nrm, ws)
    w_ref = w.double() - Vd.t() @ c.float().double()
    assert float((w1.double() - w_ref).abs().max()) <= 1e-5 * float(w_ref.abs().max())
    assert abs(nrm.item() - float(w_ref @ w_ref)) <= 1e-5 * nrm.item()
    # determinism: same inputs -> same bits (fixed-order reductions)
    c2 = torch.zeros_like(c)
    K.cgs_project(V, rows, w, c2, ws)
    assert torch.equal(c, c2)
    # sign=+1 restores w
    K.cgs_update(V, rows, c, w1, None, ws, sign=1.0)
    assert float((w1 - w).abs().max()) <= 1e-5 * float(w.abs().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,n", [(1, 1), (1, 300), (3, 2048), (8, 4096), (9, 2048 * 3 + 5), (26, 1_000_003), (50, 300_000),
                                    (51, 300_000), (100, 65536 + 17), (100, 2_000_000), (101, 70_000), (104, 70_001), (200, 40_000), (208, 3_001),
                                    (57, 255), (12, 257)])
def test_cgs_update_project_fused(K, cuda_dev, dtype, rows, n):
    """Fused middle pass of CGS2 (TMA-staged slab): w' = w - V^T c ; c2 = V w' ; |w'|^2 in one read of V,
    vs fp64 and vs the unfused update + project pair (same inputs)."""
    if rows > K.fused_max_rows(dtype):
        pytest.skip("beyond the shared-memory slab for this storage type")
    V, w = _basis(rows, n, cuda_dev, dtype, rows * 11 + n)
    ws = _ws(K, cuda_dev, rows + 1)
    g = torch.Generator(device=cuda_dev).manual_seed(rows + n)
    c = torch.randn(rows, dtype=torch.float64, device=cuda_dev, generator=g)
    w1 = w.clone()
    c2 = torch.full((rows,), float("nan"), dtype=torch.float64, device=cuda_dev)
    nrm = torch.full((1,), float("nan"), dtype=torch.float64, device=cuda_dev)
    K.cgs_update_project(V, rows, c, w1, c2, nrm, ws)
    Vd = V[:, :n].double()
    w_ref = w.double() - Vd.t() @ c.float().double()
    assert float((w1.double() - w_ref).abs().max()) <= 1e-5 * float(w_ref.abs().max()) + 1e-12
    assert abs(nrm.item() - float(w_ref @ w_ref)) <= 1e-5 * float(w_ref @ w_ref) + 1e-12
    c2_ref = Vd @ w1.double()                                    # projection of the w' the kernel produced
    tol = 3e-6 * float((Vd.abs() @ w1.double().abs()).max()) + 1e-12
    assert float((c2 - c2_ref).abs().max()) <= tol
    # unfused pair on the same inputs agrees
    w2 = w.clone()
    c2u = torch.zeros_like(c2)
    nrm_u = torch.zeros_like(nrm)
    K.cgs_update(V, rows, c, w2, nrm_u, ws)
    K.cgs_project(V, rows, w2, c2u, ws)
    assert float((w1 - w2).abs().max()) <= 2e-6 * float(w2.abs().max()) + 1e-12
    assert float((c2 - c2u).abs().max()) <= 2 * tol
    # deterministic
    w3 = w.clone()
    c3 = torch.zeros_like(c2)
    K.cgs_update_project(V, rows, c, w3, c3, nrm, ws)
    assert torch.equal(w3, w1) and torch.equal(c3, c2)


def test_cgs_update_project_limits(K, cuda_dev):
    from hessian_llm_vision_b200._lib import HLVError
    cap = K.fused_max_rows(torch.float32)
    assert cap >= 100 and K.fused_max_rows(torch.bfloat16) == 2 * cap       # m=100 fp32 must fit
    V, w = _basis(cap + 1, 4096, cuda_dev, torch.float32, 1)
    ws = _ws(K, cuda_dev, cap + 2)
    c = torch.zeros(cap + 1, dtype=torch.float64, device=cuda_dev)
    with pytest.raises(HLVError, match="exceeds the fused kernel"):
        K.cgs_update_project(V, cap + 1, c, w, c.clone(), torch.zeros(1, dtype=torch.float64, device=cuda_dev), ws)
    small = _ws(K, cuda_dev, 8)
    with pytest.raises(HLVError, match="workspace too small"):
        K.cgs_update_project(V, 8, c, w, c.clone(), torch.zeros(1, dtype=torch.float64, device=cuda_dev), small)


def test_cgs2_orthogonalises_full_size(K, cuda_dev):
    """BASELINE full size (n = GPT-2's P), size-independent properties: after CGS2 against an
    orthonormal basis the result is orthogonal to every row, and the pass is idempotent."""
    n, rows = 124_046_592, 12
    g = torch.Generator(device=cuda_dev).manual_seed(0)
    V = torch.zeros(rows, n, device=cuda_dev)
    ws = _ws(K, cuda_dev, rows)
    c = torch.zeros(rows, dtype=torch.float64, device=cuda_dev)
    nrm = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    beta = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    for j in range(rows):                       # build an orthonormal basis with the kernels themselves
        w = torch.randn(n, device=cuda_dev, generator=g)
        if j:
            for _ in range(2):
                K.cgs_project(V, j, w, c, ws)
                K.cgs_update(V, j, c, w, nrm, ws)
        else:
            K.dot(w, w, nrm, ws)
        K.normalize_store(w, nrm, beta, V[j], None)
    Vd = V.double()
    G = Vd @ Vd.t()
    del Vd
    assert float((G - torch.eye(rows, dtype=torch.float64, device=cuda_dev)).abs().max()) < 1e-6
    w = torch.randn(n, device=cuda_dev, generator=g)
    for _ in range(2):
        K.cgs_project(V, rows, w, c, ws)
        K.cgs_update(V, rows, c, w, nrm, ws)
    K.cgs_project(V, rows, w, c, ws)
    assert float(c.abs().max()) < 1e-5 * nrm.item() ** 0.5       # orthogonal to every row
    w2 = w.clone()
    K.cgs_update(V, rows, c, w2, nrm, ws)
    assert float((w2 - w).abs().max()) <= 1e-6 * float(w.abs().max())   # idempotent
    # the same CGS2 as 3 passes (fused middle pass, 2-D TMA slab) at full size: same properties, same result
    w3 = torch.randn(n, device=cuda_dev, generator=torch.Generator(device=cuda_dev).manual_seed(99))
    w4 = w3.clone()
    c2 = torch.zeros_like(c)
    ws1 = _ws(K, cuda_dev, rows + 1)
    K.cgs_project(V, rows, w3, c, ws1)
    K.cgs_update_project(V, rows, c, w3, c2, nrm, ws1)
    K.cgs_update(V, rows, c2, w3, nrm, ws1)
    for _ in range(2):
        K.cgs_project(V, rows, w4, c, ws)
        K.cgs_update(V, rows, c, w4, nrm, ws)
    assert float((w3 - w4).abs().max()) <= 2e-6 * float(w4.abs().max())
    K.cgs_project(V, rows, w3, c, ws)
    assert float(c.abs().max()) < 1e-5 * float(torch.linalg.vector_norm(w3.double()))
    assert abs(nrm.item() - float(w4.double() @ w4.double())) <= 1e-6 * nrm.item()


# ------------------------------------------------------------------ vector_adjust / Ritz vectors
@pytest.mark.parametrize("k,n", [(1, 33), (5, 700), (10, 4096), (10, 100_003)])
def test_vector_adjust_vs_oracle_and_reference_kernel(K, cuda_dev, k, n):
    from hessian_llm_vision_b200 import adjust
    from oracle import ref_native
    g = torch.Generator(device=cuda_dev).manual_seed(k * n)
    V = torch.randn(k, n, device=cuda_dev, generator=g) / n ** 0.5
    if n % 4:
        ld = (n + 7) // 8 * 8
        Vp = torch.zeros(k, ld, device=cuda_dev)
        Vp[:, :n] = V
        Vk = Vp[:, :n]
    else:
        Vk = V
    grad = torch.randn(n, device=cuda_dev, generator=g)
    eig = (torch.randn(k, device=cuda_dev, generator=g) * 5).abs() + 0.1
    delta = 1e-2
    adj = grad.clone()
    adjust.cuda_vector_adjust(grad, Vk, eig, adj, delta)
    ref = oracle.lowrank_adjust(grad.cpu(), V.cpu(), eig.cpu(), delta)
    assert float((adj.cpu() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    if n <= 4096 and ref_native.have_cubin():
        # the reference's own kernel (O(k n^2)), compiled from /root/reference/vector_adjust.cu
        adj_ref = grad.clone()
        ref_native.vector_adjust_gpu(grad, V.contiguous(), eig, adj_ref, delta)
        assert float((adj - adj_ref).abs().max()) <= 2e-5 * float(adj_ref.abs().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("m,nvec,n", [(4, 4, 100), (25, 3, 70_001), (100, 17, 200_000)])
def test_ritz_vectors(K, cuda_dev, dtype, m, nvec, n):
    Q, _ = _basis(m, n, cuda_dev, dtype, m + n)
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    Y = torch.randn(m, nvec, device=cuda_dev, generator=g)
    ld = (n + 7) // 8 * 8
    out = torch.full((nvec, ld), float("nan"), device=cuda_dev)
    K.ritz_vectors(Q, m, Y, out, n)
    ref = Y.double().t() @ Q[:, :n].double()                     # eigvects.t() @ Q
    assert float((out[:, :n].double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("m,nvec,n", [(100, 100, 128 * 300), (100, 100, 128 * 301 + 77), (104, 112, 128 * 150), (13, 9, 1000),
                                      (128, 40, 128 * 64 + 5), (50, 16, 128), (100, 113, 128 * 40), (57, 100, 100_003)])
def test_ritz_vectors_tensor_core_pass(K, cuda_dev, m, nvec, n):
    """More than 8 Ritz vectors from an fp32 basis: ONE pass over Q on tcgen05 (kind::tf32, 3xTF32 split, accumulators in
    TMEM) for the whole-128-column tiles, the CUDA-core kernel for the ragged tail -- against float64.  Y orthonormal
    (eigenvectors of T), Q orthonormal rows: every output entry is a length-m dot product, error bar 3e-6 of the scale
    (3xTF32 carries ~2^-21 per product; the CUDA-core kernel with fp32 FMAs is at ~1e-7)."""
    Q, _ = _basis(m, n, cuda_dev, torch.float32, m + n)
    g = torch.Generator(device=cuda_dev).manual_seed(m * 131 + nvec)
    Yfull, _ = torch.linalg.qr(torch.randn(m, m, device=cuda_dev, generator=g, dtype=torch.float64))
    pad = max(nvec - m, 0)                                       # nvec may exceed m in a kernel test (113 > 112: two slices)
    Y = torch.cat([Yfull, torch.randn(m, pad, device=cuda_dev, generator=g, dtype=torch.float64)], dim=1)[:, :nvec].float().contiguous()
    ld = (n + 7) // 8 * 8
    out = torch.full((nvec, ld), float("nan"), device=cuda_dev)
    K.ritz_vectors(Q, m, Y, out, n)
    ref = Y.double().t() @ Q[:, :n].double()
    scale = float(ref.abs().max())
    err = float((out[:, :n].double() - ref).abs().max())
    assert not torch.isnan(out[:, :n]).any()
    assert err <= 3e-6 * scale, (err / scale)
    assert torch.isnan(out[:, n:]).all()                         # nothing written past n
    # the same call restricted to 8 vectors goes through the CUDA-core kernel: the two paths agree
    out8 = torch.full((8, ld), float("nan"), device=cuda_dev)
    K.ritz_vectors(Q, m, Y[:, :8].contiguous(), out8, n)
    assert float((out8[:, :n] - out[:8, :n]).abs().max()) <= 3e-6 * scale


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("m,k,n", [(6, 6, 1000), (24, 5, 70_003)])
def test_adjust_implicit_equals_explicit_ritz_vectors(K, cuda_dev, dtype, m, k, n):
    """g += Q^T (Y diag(s) Y^T) (Q g) (two passes over the Lanczos basis) equals forming the k Ritz vectors
    V = Y^T Q (gpt2_hessian_cpu.py:217) and applying the reference's adjustment (:224-229) to them."""
    from hessian_llm_vision_b200 import adjust
    Q, _ = _basis(m, n, cuda_dev, dtype, 3 * m + n)
    g = torch.Generator(device=cuda_dev).manual_seed(m + n)
    A = torch.randn(m, m, device=cuda_dev, generator=g, dtype=torch.float64)
    lam, Y = torch.linalg.eigh(A + A.t())
    lam = lam.abs() + 0.5
    grad = torch.randn(n, device=cuda_dev, generator=g)
    delta = 0.05
    sel = list(range(m - k, m))                                   # the k largest, as the reference keeps them
    got = adjust.adjust_gradient_implicit(grad, Q, m, Y, lam, delta, select=sel)
    Vr = (Y[:, sel].t() @ Q[:, :n].double()).float()             # explicit Ritz vectors (fp64 product of the stored rows)
    ref = oracle.lowrank_adjust(grad.cpu(), Vr.cpu(), lam[sel].float().cpu(), delta)
    assert float((got.cpu() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    # in place, all pairs
    out = grad.clone()
    adjust.adjust_gradient_implicit(out, Q, m, Y, lam, delta, out=out)
    Va = (Y.t() @ Q[:, :n].double()).float()
    ref_all = oracle.lowrank_adjust(grad.cpu(), Va.cpu(), lam.float().cpu(), delta)
    assert float((out.cpu() - ref_all).abs().max()) <= 2e-5 * float(ref_all.abs().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conditional_update_device_predicate(K, cuda_dev, dtype):
    """hlv_cgs_needs_pass / hlv_cgs_update_if_*: the test and the predication happen on the device."""
    rows, n = 20, 70_003
    V, w = _basis(rows, n, cuda_dev, dtype, 77)
    ws = _ws(K, cuda_dev, rows)
    nrm = torch.tensor([4.0], dtype=torch.float64, device=cuda_dev)              # |w| = 2
    flag = torch.full((1,), 7, dtype=torch.int32, device=cuda_dev)
    c = torch.full((rows,), 1e-6, dtype=torch.float64, device=cuda_dev)
    K.cgs_needs_pass(c, rows, nrm, 1e-5, flag)
    assert int(flag.item()) == 0                                                 # every |c_i| <= 1e-5 * 2
    c[13] = -3e-5
    K.cgs_needs_pass(c, rows, nrm, 1e-5, flag)
    assert int(flag.item()) == 1
    K.cgs_needs_pass(c, 13, nrm, 1e-5, flag)                                     # the large one is outside the first 13 rows
    assert int(flag.item()) == 0
    c[2] = float("nan")
    K.cgs_needs_pass(c, rows, nrm, 1e30, flag)
    assert int(flag.item()) == 1                                                 # NaN asks for the pass
    # predicated update: flag 0 leaves w and norm2 untouched, flag 1 is the plain update
    c = torch.randn(rows, dtype=torch.float64, device=cuda_dev)
    w0, n0 = w.clone(), torch.tensor([123.0], dtype=torch.float64, device=cuda_dev)
    flag.zero_()
    K.cgs_update(V, rows, c, w, n0, ws, run_flag=flag)
    assert torch.equal(w, w0) and n0.item() == 123.0
    flag.fill_(1)
    K.cgs_update(V, rows, c, w, n0, ws, run_flag=flag)
    w_ref = w0.clone()
    n_ref = torch.zeros(1, dtype=torch.float64, device=cuda_dev)
    K.cgs_update(V, rows, c, w_ref, n_ref, ws)
    assert torch.equal(w, w_ref) and n0.item() == n_ref.item()
    flag.zero_()                                                                 # and the workspace ticket is still usable
    K.cgs_update(V, rows, c, w, n0, ws, run_flag=flag)
    K.cgs_update(V, rows, c, w, n0, ws)
    assert n0.item() == float(w.double() @ w.double()) or abs(n0.item() - float(w.double() @ w.double())) <= 1e-5 * n0.item()
