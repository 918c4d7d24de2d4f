"""GPU parity of every kernel family, called through the C ABI (dsvgp_b200.ops -> libdsvgp_b200.so), against the
CPU oracle (kernel assembly, its backward) or against the same op in fp64 on the CPU (GEMM, Cholesky, reductions).
Tolerances: fp64 1e-10 relative, fp32 1e-4 relative (north_star); most checks are far tighter."""
import os

import pytest
import torch

from oracle import dsvgp_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
F32, F64 = torch.float32, torch.float64


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def tol(dtype, f64=1e-10, f32=1e-4):
    return f64 if dtype == F64 else f32


@pytest.fixture(scope="module")
def ops():
    from dsvgp_b200 import ops
    return ops


def _kernel_inputs(n1, n2, d, p1, p2, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    x1, x2 = torch.rand(n1, d, generator=g, dtype=F64), torch.rand(n2, d, generator=g, dtype=F64)
    v1 = torch.randn(n1 * p1, d, generator=g, dtype=F64) if p1 else None
    v2 = torch.randn(n2 * p2, d, generator=g, dtype=F64) if p2 else None
    c = lambda t: None if t is None else t.to(dtype)
    return c(x1), c(x2), c(v1), c(v2), torch.tensor([[0.3]], dtype=dtype), torch.tensor([-0.4], dtype=dtype)


def _gpu_kernel(ops, x1, x2, v1, v2, raw_ell, raw_os, k_dtype, diag_add=0.0, use_os=True):
    dev = "cuda"
    n1, n2 = x1.shape[0], x2.shape[0]
    p1 = 0 if v1 is None else v1.shape[0] // n1
    p2 = 0 if v2 is None else v2.shape[0] // n2
    hyp = ops.hyp_from_raw(raw_ell.reshape(-1).to(dev), raw_os.to(dev))
    u1 = ops.normalize_dirs(v1.to(dev), k_dtype)[0] if p1 else None
    w2 = ops.normalize_dirs(v2.to(dev), k_dtype)[0] if p2 else None
    K = torch.full((n1 * (p1 + 1) + 3, n2 * (p2 + 1) + 5), -7.0, dtype=k_dtype, device=dev)   # padded ld, sentinel
    ops.kdir_fwd(x1.to(dev).contiguous(), u1, p1, x2.to(dev).contiguous(), w2, p2, hyp, K, use_os=use_os, diag_add=diag_add)
    torch.cuda.synchronize()
    assert float(K[n1 * (p1 + 1):].max()) == -7.0 and float(K[:, n2 * (p2 + 1):].max()) == -7.0   # no out-of-bounds writes
    return K[: n1 * (p1 + 1), : n2 * (p2 + 1)]


@pytest.mark.parametrize("dtype,k_dtype", [(F64, F64), (F32, F32), (F32, F64)])
@pytest.mark.parametrize("n1,n2,d,p1,p2", [
    (7, 9, 2, 2, 2), (60, 200, 2, 2, 2), (33, 65, 3, 1, 1), (50, 70, 10, 2, 2), (40, 45, 60, 3, 3),
    (37, 129, 18, 2, 0), (20, 31, 4, 1, 0), (19, 23, 5, 3, 0), (16, 16, 7, 0, 0),
    (9, 11, 3, 4, 4), (6, 8, 5, 5, 5), (5, 7, 4, 2, 1), (1, 1, 1, 1, 1),
    (40, 600, 10, 2, 2), (17, 1000, 3, 1, 1), (33, 515, 18, 2, 0), (9, 777, 5, 3, 0), (20, 640, 4, 1, 0), (8, 512, 60, 2, 2)])
def test_kdir_forward_matches_oracle(ops, dtype, k_dtype, n1, n2, d, p1, p2):
    x1, x2, v1, v2, raw_ell, raw_os = _kernel_inputs(n1, n2, d, p1, p2, dtype, 100 + n1 + d)
    K = _gpu_kernel(ops, x1, x2, v1, v2, raw_ell, raw_os, k_dtype)
    up = lambda t: None if t is None else t.double()
    ell, osc = torch.nn.functional.softplus(raw_ell.double()).reshape(()), torch.nn.functional.softplus(raw_os.double())
    Kref = osc * O.kernel_closed_form(up(x1), up(x2), up(v1), up(v2), ell)
    assert K.shape == Kref.shape
    assert rel(K, Kref) < (1e-12 if k_dtype == F64 and dtype == F64 else 2e-6)


def test_kdir_forward_matches_reference_golden(ops):
    cases = torch.load(os.path.join(GOLD, "kernel_cases.pt"))
    for name, c in cases.items():
        dt = c["K"].dtype
        K = _gpu_kernel(ops, c["x1"], c["x2"], c["v1"], c["v2"], c["raw_ell"], torch.zeros(1, dtype=dt), dt, use_os=False)
        assert rel(K, c["K"]) < (1e-12 if dt == F64 else 2e-6), name
        if "Kdiag" in c:
            hyp = ops.hyp_from_raw(c["raw_ell"].reshape(-1).cuda())
            n, p = c["x1"].shape[0], c["v1"].shape[0] // c["x1"].shape[0]
            assert rel(ops.kdir_diag(n, p, hyp, dt, use_os=False), c["Kdiag"]) < 1e-12


@pytest.mark.parametrize("n1,n2,d,p1,p2", [(40, 600, 10, 2, 2), (17, 1000, 3, 1, 1), (70, 515, 18, 3, 3), (8, 640, 60, 2, 2)])
def test_kdir_forward_canonical_fast_path(ops, n1, n2, d, p1, p2):
    """One-hot data-side directions (any sign / scale), detected on the device, must give the same matrix as the
    general path; a single non-canonical row must switch the whole call back to the general path."""
    x1, x2, v1, _, raw_ell, raw_os = _kernel_inputs(n1, n2, d, p1, p2, F32, 300 + n1)
    g = torch.Generator().manual_seed(n2)
    idx = torch.randint(0, d, (n2 * p2,), generator=g)
    scale = torch.randn(n2 * p2, generator=g)
    scale[scale.abs() < 0.1] = 0.7
    v2 = torch.zeros(n2 * p2, d)
    v2[torch.arange(n2 * p2), idx] = scale
    for make_general in (False, True):
        if make_general:
            v2[5, (idx[5] + 1) % d] = 0.3
        hyp = ops.hyp_from_raw(raw_ell.reshape(-1).cuda(), raw_os.cuda())
        u1 = ops.normalize_dirs(v1.cuda())[0]
        w2, _, cidx, flag = ops.normalize_dirs_canon(v2.cuda())
        assert int(flag.item()) == (0 if make_general else 1)
        K = torch.full((n1 * (p1 + 1), (n2 * (p2 + 1) + 31) // 32 * 32), -7.0, device="cuda")
        ops.kdir_fwd(x1.cuda(), u1, p1, x2.cuda(), w2, p2, hyp, K, canon=(cidx, flag))
        ell, osc = torch.nn.functional.softplus(raw_ell.double()).reshape(()), torch.nn.functional.softplus(raw_os.double())
        Kref = osc * O.kernel_closed_form(x1.double(), x2.double(), v1.double(), v2.double(), ell)
        assert rel(K[:, : n2 * (p2 + 1)], Kref) < 2e-6
        assert float(K[:, n2 * (p2 + 1):].max()) == -7.0 if K.shape[1] > n2 * (p2 + 1) else True


def test_kdir_jitter_and_symmetry(ops):
    x1, _, v1, _, raw_ell, raw_os = _kernel_inputs(70, 1, 6, 2, 0, F64, 5)
    K = _gpu_kernel(ops, x1, x1, v1, v1, raw_ell, raw_os, F64, diag_add=1e-3)
    K0 = _gpu_kernel(ops, x1, x1, v1, v1, raw_ell, raw_os, F64)
    assert rel(K - K0, 1e-3 * torch.eye(K.shape[0], dtype=F64)) < 1e-9
    assert rel(K, K.T) < 1e-13
    assert torch.linalg.eigvalsh(K.cpu()).min() > 0


@pytest.mark.parametrize("dtype,k_dtype", [(F64, F64), (F32, F32), (F32, F64)])
@pytest.mark.parametrize("n1,n2,d,p1,p2", [
    (20, 50, 2, 2, 2), (33, 200, 3, 1, 1), (50, 300, 10, 2, 2), (17, 40, 60, 3, 3), (37, 129, 18, 2, 0),
    (19, 23, 5, 3, 0), (16, 16, 7, 0, 0), (9, 11, 3, 4, 4), (5, 7, 4, 2, 1), (70, 2100, 4, 2, 2),
    (100, 700, 10, 2, 2), (65, 513, 3, 1, 1), (33, 1111, 16, 2, 0), (9, 640, 7, 1, 0), (130, 900, 13, 2, 2)])
def test_kdir_backward_matches_oracle_autograd(ops, dtype, k_dtype, n1, n2, d, p1, p2):
    _check_kdir_backward(ops, dtype, k_dtype, n1, n2, d, p1, p2, n2 * (p2 + 1) + 3)


@pytest.mark.parametrize("vpl", [4, 2])
@pytest.mark.parametrize("n1,n2,d,p1,p2", [
    (100, 1024, 10, 2, 2), (70, 1000, 10, 2, 2), (130, 640, 18, 2, 0), (65, 768, 3, 1, 1), (20, 1300, 18, 1, 0),
    (40, 512, 10, 2, 0), (9, 896, 12, 2, 2)])
def test_kdir_backward_staged_upstream_rows(ops, vpl, n1, n2, d, p1, p2):
    """the layout the training step uses (16-byte aligned rows, leading dimension a multiple of 64): whole column tiles take the
    cp.async-staged upstream rows, the ragged last tile the direct loads; exact-d instantiations (d = 10, 18) beside padded ones"""
    ops.set_kdir_bwd_vpl(vpl)
    try:
        _check_kdir_backward(ops, F32, F32, n1, n2, d, p1, p2, (n2 * (p2 + 1) + 63) // 64 * 64, second=False)
    finally:
        ops.set_kdir_bwd_vpl(4)


def _check_kdir_backward(ops, dtype, k_dtype, n1, n2, d, p1, p2, ld, second=True):
    x1, x2, v1, v2, raw_ell, raw_os = _kernel_inputs(n1, n2, d, p1, p2, dtype, 200 + n1 + d)
    g = torch.Generator().manual_seed(7)
    dK = torch.randn(n1 * (p1 + 1), n2 * (p2 + 1), generator=g, dtype=F64)
    # oracle gradients (fp64 autograd through the closed form, incl. the normalisation)
    up = lambda t: None if t is None else t.double().clone().requires_grad_(True)
    X1, X2, V1, V2 = up(x1), up(x2), up(v1), up(v2)
    ell = torch.nn.functional.softplus(raw_ell.double()).reshape(()).clone().requires_grad_(True)
    osc = torch.nn.functional.softplus(raw_os.double()).clone().requires_grad_(True)
    (osc * O.kernel_closed_form(X1, X2, V1, V2, ell) * dK).sum().backward()

    dev = "cuda"
    hyp = ops.hyp_from_raw(raw_ell.reshape(-1).to(dev), raw_os.to(dev))
    u1, inv1 = ops.normalize_dirs(v1.to(dev), k_dtype) if p1 else (None, None)
    w2, inv2 = ops.normalize_dirs(v2.to(dev), k_dtype) if p2 else (None, None)
    dKd = torch.zeros(n1 * (p1 + 1), ld, dtype=k_dtype, device=dev)
    dKd[:, : n2 * (p2 + 1)] = dK.to(k_dtype)
    dKd = dKd[:, : n2 * (p2 + 1)]
    z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
    gx1, gv1, gsc = z(n1, d), (z(n1 * p1, d) if p1 else None), z(2)
    ops.kdir_bwd(x1.to(dev), u1, inv1, p1, x2.to(dev), w2, p2, hyp, dKd, gx1, gv1, gsc)
    t = 1e-11 if (dtype == F64) else 2e-5
    assert rel(gx1, X1.grad) < t
    if p1:
        assert rel(gv1, V1.grad) < t
    if second:
        gx2, gv2 = z(n2, d), (z(n2 * p2, d) if p2 else None)
        ops.kdir_bwd(x2.to(dev), w2, inv2, p2, x1.to(dev), u1, p1, hyp, dKd, gx2, gv2, None, dk_trans=True)
        assert rel(gx2, X2.grad) < t
        if p2:
            assert rel(gv2, V2.grad) < t
    assert abs(float(gsc[0]) - float(ell.grad)) < t * max(1.0, abs(float(ell.grad)))
    assert abs(float(gsc[1]) - float(osc.grad)) < t * max(1.0, abs(float(osc.grad)))


def test_kernel_module_autograd_matches_oracle():
    """RBFKernelDirectionalGrad.forward(x1, x2, v1=, v2=) as a differentiable torch op."""
    from dsvgp_b200 import gp
    x1, x2, v1, v2, raw_ell, _ = _kernel_inputs(12, 15, 4, 2, 2, F64, 9)
    k = gp.RBFKernelDirectionalGrad().to("cuda", F64)
    k.raw_lengthscale.data.copy_(raw_ell)
    dv = lambda t: t.cuda().requires_grad_(True)
    a, b, c, e = dv(x1), dv(x2), dv(v1), dv(v2)
    K = k(a, b, v1=c, v2=e)
    w = torch.randn(K.shape, dtype=F64, generator=torch.Generator().manual_seed(1)).cuda()
    (K * w).sum().backward()
    A, B, C, E = (t.clone().requires_grad_(True) for t in (x1, x2, v1, v2))
    re = raw_ell.clone().requires_grad_(True)
    Kr = O.kernel_closed_form(A, B, C, E, torch.nn.functional.softplus(re).reshape(()))
    (Kr * w.cpu()).sum().backward()
    assert rel(K, Kr) < 1e-12
    for got, want in ((a.grad, A.grad), (b.grad, B.grad), (c.grad, C.grad), (e.grad, E.grad), (k.raw_lengthscale.grad, re.grad)):
        assert rel(got, want) < 1e-10
    with pytest.raises(AssertionError):
        k(a, b, v1=c, v2=e[:15])
    with pytest.raises(RuntimeError, match="diag=True only works"):
        k(a, b, diag=True, v1=c, v2=e)
    assert rel(k(a, a, diag=True, v1=c, v2=c), O.kernel_diag(12, 2, torch.nn.functional.softplus(raw_ell).reshape(()))) < 1e-12
    assert k.num_outputs_per_input(a, b) == 3


# ------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (70, 130, 45), (1, 1, 1), (200, 37, 300), (129, 257, 64)])
def test_gemm_dense(ops, dtype, ta, tb, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g, dtype=F64)
    B = torch.randn((N, K) if tb else (K, N), generator=g, dtype=F64)
    C0 = torch.randn(M, N, generator=g, dtype=F64)
    ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 1.3 * C0
    C = C0.to(dtype).cuda()
    ops.gemm(A.to(dtype).cuda(), B.to(dtype).cuda(), C, ta=ta, tb=tb, alpha=0.7, beta=-1.3)
    assert rel(C, ref) < tol(dtype, 1e-13, 5e-6)


@pytest.mark.parametrize("dtype", [F64, F32])
def test_gemm_triangular_flags_and_addend(ops, dtype):
    g = torch.Generator().manual_seed(3)
    n, m = 150, 210
    Lr = torch.randn(n, n, generator=g, dtype=F64)           # garbage above the diagonal must be ignored
    L = Lr.tril()
    B = torch.randn(n, m, generator=g, dtype=F64)
    D = torch.randn(n, m, generator=g, dtype=F64)
    d = lambda t: t.to(dtype).cuda()
    t = tol(dtype, 1e-13, 2e-6)
    C = torch.empty(n, m, dtype=dtype, device="cuda")
    ops.gemm(d(Lr), d(B), C, a_tri=ops.TRI_LOWER)
    assert rel(C, L @ B) < t
    ops.gemm(d(Lr), d(B), C, ta=True, a_tri=ops.TRI_UPPER)
    assert rel(C, L.T @ B) < t
    ops.gemm(d(Lr), d(B), C, a_tri=ops.TRI_LOWER, alpha=2.0, beta=-2.0, D=d(D))
    assert rel(C, 2 * L @ B - 2 * D) < t
    C2 = torch.empty(n, m, dtype=dtype, device="cuda")
    ops.gemm(d(Lr), d(B), C, ta=True, a_tri=ops.TRI_UPPER, C2=C2, D2=d(D))        # second output C2 = C + D2
    assert rel(C, L.T @ B) < t and rel(C2, L.T @ B + D) < t
    E = torch.empty(n, n, dtype=dtype, device="cuda")
    ops.tril_minus_eye(d(Lr), E)
    assert rel(E, L - torch.eye(n, dtype=F64)) < 1e-7
    S = torch.empty(n, n, dtype=dtype, device="cuda")
    R = torch.randn(n, n, generator=g, dtype=F64)
    ops.gemm(d(R), d(Lr), S, b_tri=ops.TRI_LOWER)
    assert rel(S, R @ L) < t
    ops.gemm(d(R), d(Lr), S, tb=True, b_tri=ops.TRI_UPPER)
    assert rel(S, R @ L.T) < t
    # SYRK-style: lower tiles only, then mirror
    S.fill_(float("nan"))
    ops.gemm(d(B), d(B), S, tb=True, c_tri=1)
    ops.mirror_lower(S)
    assert rel(S, B @ B.T) < tol(dtype, 1e-13, 5e-6)
    # leading-dimension / sub-matrix use
    big = torch.zeros(n + 10, m + 6, dtype=dtype, device="cuda")
    ops.gemm(d(Lr), d(B), big, a_tri=ops.TRI_LOWER, M=n, N=m, K=n)
    assert rel(big[:n, :m], L @ B) < t and float(big[n:].abs().max()) == 0 and float(big[:, m:].abs().max()) == 0


# --------------------------------------------------------------------------------------------- Cholesky
@pytest.mark.parametrize("Mq", [1, 7, 60, 112, 113, 250, 1024, 1500])
def test_cholesky_and_inverse(ops, Mq):
    g = torch.Generator().manual_seed(Mq)
    R = torch.randn(Mq, Mq + 5, generator=g, dtype=F64)
    A = R @ R.T / (Mq + 5) + 1e-3 * torch.eye(Mq, dtype=F64)
    Mp, nb0, nlev = ops.chol_plan(Mq)
    Aw = torch.full((Mp, Mp), float("nan"), dtype=F64, device="cuda")
    Aw[:Mq, :Mq] = A.cuda()
    ops.pad_identity(Aw, Mq)
    L = torch.full((Mp, Mp), float("nan"), dtype=F64, device="cuda")
    W = torch.full((Mp, Mp), float("nan"), dtype=F64, device="cuda")
    info = torch.ones(1, dtype=torch.int32, device="cuda")
    ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
    assert int(info.item()) == 0
    Lr = torch.linalg.cholesky(A)
    assert rel(L[:Mq, :Mq].tril(), Lr) < 1e-11
    assert rel(W[:Mq, :Mq].tril(), torch.linalg.inv(Lr)) < 1e-9
    assert rel(W[:Mq, :Mq].tril().cpu() @ Lr, torch.eye(Mq, dtype=F64)) < 1e-10


DEFAULT_INV_STREAMS = 3


@pytest.mark.parametrize("knob", ["priority", "lookahead", "rank_update", "priority+lookahead+rank_update", "two_inverse_streams", "per_level_inverse_streams", "one_inverse_stream",
                                  "own_graph"])
def test_cholesky_scheduling_knobs_keep_the_result(ops, knob):
    """The experimental schedules of the factorisation (high-priority streams, lookahead-2 split of the trailing update, rank-K
    update kernel; all off by default) give the same factor and inverse."""
    Mq = 3072
    g = torch.Generator().manual_seed(11)
    R = torch.randn(Mq, Mq + 5, generator=g, dtype=F64)
    A = (R @ R.T / (Mq + 5) + 1e-3 * torch.eye(Mq, dtype=F64)).cuda()
    Mp, nb0, nlev = ops.chol_plan(Mq)

    def run():
        Aw = A.clone()
        L, W = (torch.full((Mp, Mp), float("nan"), dtype=F64, device="cuda") for _ in range(2))
        info = torch.ones(1, dtype=torch.int32, device="cuda")
        ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
        assert int(info.item()) == 0
        return L.tril(), W.tril()
    L0, W0 = run()
    try:
        if "priority" in knob:
            ops.set_chol_priority(1)
        if "lookahead" in knob:
            ops.set_chol_lookahead(1)
        if "rank_update" in knob:
            ops.set_rank_update(1)
        if "inverse_stream" in knob:
            ops.set_chol_inv_streams({"two": 2, "per": 3, "one": 1}[knob[:3]])
        if knob == "own_graph":
            ops.set_chol_graph(1)
        L1, W1 = run()
        if "inverse_stream" in knob or knob == "own_graph":   # same kernels in another stream layout / replayed from the cached graph: bit-identical, run after run
            for _ in range(5):
                L2, W2 = run()
                assert torch.equal(L2, L0) and torch.equal(W2, W0)
            assert torch.equal(L1, L0) and torch.equal(W1, W0)
    finally:
        ops.set_chol_priority(0), ops.set_chol_lookahead(0), ops.set_rank_update(0), ops.set_chol_inv_streams(DEFAULT_INV_STREAMS)
        ops.set_chol_graph(0)
    assert rel(L1, L0) < 1e-12 and rel(W1, W0) < 1e-10
    assert rel(W0 @ L0, torch.eye(Mp, dtype=F64)) < 1e-9


def test_limited_pairs_and_trace_do_not_change_the_product(ops):
    """dsvgp_set_tc_max_pairs (a persistent product on part of the GPU) and dsvgp_set_tc_trace (per-item clock stamps) are
    scheduling / profiling aids: the product is bit-identical, and every traced item carries increasing stamps."""
    from dsvgp_b200 import _lib
    F16 = torch.float16
    M, N, K = 768, 4096, 768
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(M, K, device="cuda", dtype=F64, generator=g).tril()
    B = torch.randn(K, N, device="cuda", dtype=F64, generator=g)
    Ah, Al = _split_ref(A, 2.0 ** 11)
    Bh, Bl = _split_ref(B, 2.0 ** 12)
    inv = torch.tensor([2.0 ** -23], device="cuda", dtype=F32)

    def run():
        C = torch.full((M, N), float("nan"), device="cuda", dtype=F32)
        ops.gemm_tch((Ah, Al), (Bh, Bl), C, M, N, K, inv, a_tri=1)
        return C
    C0 = run()
    nitems = 3 * 16
    tr = torch.zeros(8 * nitems, dtype=torch.int64, device="cuda")
    try:
        ops.set_tc_max_pairs(5)
        _lib.call_raw("dsvgp_set_tc_trace", tr, nitems)
        C1 = run()
    finally:
        ops.set_tc_max_pairs(0)
        _lib.call_raw("dsvgp_set_tc_trace", None, 0)
    assert torch.equal(C0, C1)
    t = tr.view(nitems, 8).cpu()
    assert bool((t[:, 0] > 0).all()) and bool((t[:, 1] >= t[:, 0]).all()) and bool((t[:, 2] >= t[:, 1]).all())
    assert bool((t[:, 4] >= t[:, 3]).all()) and bool((t[:, 5] >= t[:, 4]).all()) and bool((t[:, 6] > t[:, 5]).all())


def test_collect_grads_kernel(ops):
    """The end-of-step gather of every small gradient: layout, dtype cast and softplus chain rule."""
    nZ, nV = 70, 140
    g = torch.Generator().manual_seed(2)
    small = torch.randn(8 + nZ + nV, generator=g, dtype=F64).cuda()
    hyp = torch.rand(8, generator=g, dtype=F64).cuda()
    for dtype in (F32, F64):
        for mode in (0, 1):
            out = torch.full((nZ + nV + 4,), float("nan"), dtype=dtype, device="cuda")
            ops.collect_grads(small, nZ, nV, hyp, mode, out)
            want = torch.cat([small[8:], small[7:8], small[5:6] * hyp[5], small[4:5] * hyp[4],
                              ((small[1:2] if mode else 0.0) + small[6:7]) * hyp[6]])
            assert rel(out, want) < (1e-15 if dtype == F64 else 1e-7)


def test_cholesky_reports_non_positive_pivot(ops):
    A = torch.eye(8, dtype=F64)
    A[5, 5] = -1.0
    Mp, nb0, nlev = ops.chol_plan(8)
    Aw = A.cuda()
    L, W = torch.empty_like(Aw), torch.empty_like(Aw)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
    assert int(info.item()) == 6


# ------------------------------------------------------------------------------------ reductions / elementwise
@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("rows,nq", [(60, 600), (1024, 1000), (300, 5000), (3, 2)])
def test_mean_variance_reductions(ops, dtype, rows, nq):
    g = torch.Generator().manual_seed(rows + nq)
    ld = (nq + 7) // 8 * 8
    A = torch.randn(rows, ld, generator=g, dtype=F64)
    C = torch.randn(rows, ld, generator=g, dtype=F64)
    m = torch.randn(rows, generator=g, dtype=F64)
    raw = [torch.tensor([v], dtype=dtype, device="cuda") for v in (0.2, -0.3, -1.0, 0.05)]
    hyp = ops.hyp_from_raw(*raw)
    ell, osc, s2, c = (float(v) for v in hyp[:4].cpu())
    d = lambda t: t.to(dtype).cuda()
    ns = ops.reduce_slabs(rows, nq)
    pm, pv = torch.empty(ns, nq, dtype=dtype, device="cuda"), torch.empty(ns, nq, dtype=dtype, device="cuda")
    mu, var = torch.empty(nq, dtype=dtype, device="cuda"), torch.empty(nq, dtype=dtype, device="cuda")
    for p2 in (0, 2):
        if nq % (p2 + 1):
            continue
        kd = torch.tensor([osc] + [osc / ell ** 2] * p2, dtype=F64).repeat(nq // (p2 + 1))
        ops.col_dots(d(A), d(m), pm, pv, rows, nq, C=d(C))
        ops.predict_finish(pm, pv, nq, p2, hyp, mu, var, add_noise=True)
        assert rel(mu, A[:, :nq].T @ m + c) < tol(dtype, 1e-12, 1e-5)
        assert rel(var, (kd + 1e-4 + s2 + (A * C)[:, :nq].sum(0)).clamp_min(1e-10 if dtype == F64 else 1e-6)) < tol(dtype, 1e-12, 2e-5)
        ops.col_dots(d(A), d(m), pm, pv, rows, nq, Bp=d(C))       # C plays B' = B - A: sum B^2 - A^2 = sum B'(2A + B')
        ops.predict_finish(pm, pv, nq, p2, hyp, mu, var, add_noise=False)
        assert rel(var, (kd + 1e-4 + (C * (2 * A + C))[:, :nq].sum(0)).clamp_min(1e-10 if dtype == F64 else 1e-6)) < tol(dtype, 1e-12, 2e-5)


@pytest.mark.parametrize("dtype", [F64, F32])
def test_elbo_terms_kl_and_small_kernels(ops, dtype):
    g = torch.Generator().manual_seed(11)
    nq, Mq, p2 = 1000, 90, 1
    mu, y = torch.randn(nq, generator=g, dtype=F64), torch.randn(nq, generator=g, dtype=F64)
    var = torch.rand(nq, generator=g, dtype=F64) + 0.1
    raw = [torch.tensor([v], dtype=dtype, device="cuda") for v in (0.2, -0.3, -1.0, 0.05)]
    hyp = ops.hyp_from_raw(*raw)
    ell, osc, s2, c = (float(v) for v in hyp[:4].cpu())
    d = lambda t: t.to(dtype).cuda()
    gmu, gvar = torch.empty(nq, dtype=dtype, device="cuda"), torch.empty(nq, dtype=dtype, device="cuda")
    sc, ws = torch.zeros(8, dtype=F64, device="cuda"), torch.empty(8192, dtype=F64, device="cuda")
    w = 1.0 / 3000
    ops.elbo_terms(d(mu), d(var), d(y), hyp, w, gmu, gvar, sc, ws)
    MU, VAR, S2 = mu.clone().requires_grad_(True), var.clone().requires_grad_(True), torch.tensor(s2, dtype=F64, requires_grad=True)
    val = (-0.5 * (((y - MU) ** 2 + VAR) / S2 + S2.log() + torch.log(torch.tensor(2 * torch.pi, dtype=F64))) * w).sum()
    val.backward()
    t = tol(dtype, 1e-12, 1e-5)
    assert abs(float(sc[0]) - float(val)) < t * abs(float(val))
    assert abs(float(sc[1]) - float(S2.grad)) < t * abs(float(S2.grad))
    assert rel(gmu, MU.grad) < t and rel(gvar, VAR.grad) < t
    gs = torch.zeros(4, dtype=F64, device="cuda")
    ops.pred_bwd_scalars(gmu, gvar, p2, hyp, True, gs, ws)
    gv = VAR.grad
    want = [(gv[1::2] * (-2 * osc / ell ** 3)).sum(), (gv[0::2]).sum() + (gv[1::2] / ell ** 2).sum(), gv.sum(), MU.grad.sum()]
    for k in range(4):
        assert abs(float(gs[k]) - float(want[k])) < t * max(1e-3, abs(float(want[k]))), k
    # KL and the variational-parameter gradients
    m = torch.randn(Mq, generator=g, dtype=F64)
    Ls = torch.randn(Mq, Mq, generator=g, dtype=F64) * 0.1 + torch.eye(Mq, dtype=F64)
    H = torch.randn(Mq, Mq, generator=g, dtype=F64)
    tt = torch.randn(Mq, generator=g, dtype=F64)
    kl = torch.zeros(1, dtype=F64, device="cuda")
    ops.kl_divergence(d(m), d(Ls), kl, ws)
    Pm, PL = m.clone().requires_grad_(True), Ls.clone().requires_grad_(True)
    P = O.Params(Z=None, Vz=None, m=Pm, Ls_raw=PL, c=None, raw_os=None, raw_ell=None, raw_noise=None)
    klr = O.kl_divergence(P)
    klr.backward()
    assert abs(float(kl) - float(klr)) < t * abs(float(klr))
    gm, gLs = torch.empty(Mq, dtype=dtype, device="cuda"), torch.empty(Mq, Mq, dtype=dtype, device="cuda")
    ops.var_grads(d(H), d(Ls), d(tt), d(m), 1.0 / 777, gm, gLs)
    assert rel(gm, tt - Pm.grad / 777) < t
    assert rel(gLs, (2 * H.T).tril() - PL.grad / 777) < t
    # dA / t / Ag
    rows, nq2 = 70, 333
    A, C = torch.randn(rows, 336, generator=g, dtype=F64), torch.randn(rows, 336, generator=g, dtype=F64)
    mm, gm2, gv2 = torch.randn(rows, generator=g, dtype=F64), torch.randn(nq2, generator=g, dtype=F64), torch.randn(nq2, generator=g, dtype=F64)
    Cd, Ag = d(C), torch.zeros(rows, 336, dtype=dtype, device="cuda")
    tp, tv = torch.empty(4, rows, dtype=dtype, device="cuda"), torch.empty(rows, dtype=dtype, device="cuda")
    ops.dA_apply(d(A), Cd, Ag, rows, nq2, d(mm), d(gm2), d(gv2), tp, tv)
    assert rel(Cd[:, :nq2], mm[:, None] * gm2[None] + 2 * C[:, :nq2] * gv2[None]) < t
    assert rel(Ag[:, :nq2], A[:, :nq2] * gv2[None]) < t
    assert rel(tv, A[:, :nq2] @ gm2) < tol(dtype, 1e-12, 2e-5)
    # sym_phi, add_outer, casts
    Y = torch.randn(50, 50, generator=g, dtype=F64)
    Pp = torch.empty(50, 50, dtype=F64, device="cuda")
    ops.sym_phi(Y.cuda(), Pp, 50)
    Phi = Y.tril()
    Phi.diagonal().mul_(0.5)
    assert rel(Pp, 0.5 * (Phi + Phi.T)) < 1e-14
    X = d(Y)
    ops.add_outer(X, d(m[:50]), d(tt[:50]), 1.0)
    assert rel(X, Y + m[:50, None] * tt[None, :50]) < t
    dst = torch.empty(50, 50, dtype=F32 if dtype == F64 else F64, device="cuda")
    ops.cast2d(d(Y), dst, tril=True)
    assert rel(dst, Y.tril()) < 1e-6


# ------------------------------------------------------------------------- 3xFP16 tensor-core product and its operands
def _split_ref(x, s):
    xs = x.float() * s
    hi = xs.half()
    return hi, (xs - hi.float()).half()


def _padded(t, ld):
    out = torch.zeros(t.shape[0], ld, device=t.device, dtype=t.dtype)
    out[:, :t.shape[1]] = t
    return out[:, :t.shape[1]]


@pytest.mark.parametrize("cg,persistent", [(1, 1), (2, 1), (2, 0)])
@pytest.mark.parametrize("M,N,K,kw", [
    (128, 256, 64, {}), (200, 300, 100, {}), (384, 1000, 384, dict(a_tri=1)), (384, 1000, 384, dict(a_tri=2)),
    (384, 1000, 384, dict(a_tri=1, alpha=2.0, beta=-2.0, dual=True)), (300, 300, 1000, dict(b_kmajor=True)),
    (512, 512, 2048, dict(b_kmajor=True, c_lower=True)), (512, 512, 8192, dict(b_kmajor=True, c_lower=True, nsplit=4)),
    (256, 700, 512, dict(chunk=2)), (3200, 2100, 3200, dict(a_tri=1, dual=True)), (1536, 1536, 40000, dict(b_kmajor=True, c_lower=True, nsplit=8))])
def test_gemm_tch_matches_fp64(ops, cg, persistent, M, N, K, kw):
    """tcgen05 kind::f16 x3 on scaled two-half operands against an fp64 product of the SAME quantised operands (so the
    tolerance measures the kernel: fp32 accumulation + the dropped lo*lo term), every epilogue output included.
    (cg, persistent): one CTA per tile; CTA pairs as persistent pairs over a work list (the default -- stream-K for the split-K
    cases, more items than pairs for the last two shapes); CTA pairs, one pair per tile."""
    F16 = torch.float16
    b_kmajor, a_tri, c_lower = kw.get("b_kmajor", False), kw.get("a_tri", 0), kw.get("c_lower", False)
    alpha, beta, dual, nsplit, chunk = kw.get("alpha", 1.0), kw.get("beta", 0.0), kw.get("dual", False), kw.get("nsplit", 1), kw.get("chunk", 1)
    prev = ops.set_tc_cta_group(cg)
    ops.set_tc_persistent(persistent)
    try:
        g = torch.Generator(device="cuda").manual_seed(M + N + K)
        A = torch.randn(M, K, device="cuda", dtype=F64, generator=g)
        A = A.tril() if a_tri == 1 else A.triu() if a_tri == 2 else A
        B = torch.randn((N, K) if b_kmajor else (K, N), device="cuda", dtype=F64, generator=g)
        sA, sB = 2.0 ** 11, 2.0 ** 12
        lda, ldn = (K + 7) // 8 * 8, (N + 63) // 64 * 64
        Ah, Al = (_padded(t, lda) for t in _split_ref(A, sA))
        Bh, Bl = (_padded(t, lda if b_kmajor else ldn) for t in _split_ref(B, sB))
        D, D2 = (_padded(torch.randn(M, N, device="cuda", dtype=F32, generator=g), ldn) for _ in range(2))
        C, C2 = (_padded(torch.full((M, N), float("nan"), device="cuda", dtype=F32), ldn) for _ in range(2))
        Ch, C2h = (tuple(_padded(torch.zeros(M, N, device="cuda", dtype=F16), ldn) for _ in range(2)) for _ in range(2))
        inv = torch.tensor([1.0 / (sA * sB)], device="cuda", dtype=F32)
        cs, c2s = torch.tensor([8.0], device="cuda"), torch.tensor([4.0], device="cuda")
        Aq, Bq = (Ah.double() + Al.double()) / sA, (Bh.double() + Bl.double()) / sB
        ref = alpha * (Aq @ (Bq.T if b_kmajor else Bq)) + beta * D.double()
        ws = torch.empty(nsplit * M * ((N + 3) // 4 * 4), device="cuda") if nsplit > 1 else None
        ops.gemm_tch((Ah, Al), (Bh, Bl), C, M, N, K, inv, b_kmajor=b_kmajor, alpha=alpha, beta=beta, D=D if beta else None,
                     C2=C2 if dual else None, D2=D2 if dual else None, Ch=Ch if dual else None, c_scale=cs,
                     C2h=C2h if dual else None, c2_scale=c2s, a_tri=a_tri, c_lower=c_lower, chunk=chunk, nsplit=nsplit, split_ws=ws)
        assert (rel(C.tril(), ref.tril()) if c_lower else rel(C, ref)) < 2e-6
        if dual:
            assert rel(C2, ref + D2.double()) < 2e-6
            assert rel((Ch[0].double() + Ch[1].double()) / 8.0, ref) < 2e-6
            assert rel((C2h[0].double() + C2h[1].double()) / 4.0, ref + D2.double()) < 2e-6
        if cg == 2 and persistent and nsplit == 1:
            # the persistent kernel adds the same chunks in the same order: bit-identical to one pair per tile
            Cp = C.clone()
            C.fill_(float("nan"))
            ops.set_tc_persistent(0)
            ops.gemm_tch((Ah, Al), (Bh, Bl), C, M, N, K, inv, b_kmajor=b_kmajor, alpha=alpha, beta=beta, D=D if beta else None,
                         C2=C2 if dual else None, D2=D2 if dual else None, Ch=Ch if dual else None, c_scale=cs,
                         C2h=C2h if dual else None, c2_scale=c2s, a_tri=a_tri, c_lower=c_lower, chunk=chunk, nsplit=nsplit, split_ws=ws)
            a_, b_ = (Cp.tril(), C.tril()) if c_lower else (Cp, C)
            assert torch.equal(a_.contiguous().view(torch.int32), b_.contiguous().view(torch.int32))
    finally:
        ops.set_tc_cta_group(prev)
        ops.set_tc_persistent(1)


@pytest.mark.parametrize("cg,persistent", [(1, 1), (2, 1), (2, 0)])
@pytest.mark.parametrize("M,N,K,kw", [
    (128, 256, 64, {}), (200, 300, 100, {}), (384, 1000, 384, dict(a_tri=1, alpha=2.0, beta=2.0)),
    (384, 1000, 384, dict(a_tri=2, dual=True)), (512, 512, 512, dict(b_kmajor=True, a_tri=1, c_lower=True)),
    (600, 600, 4096, dict(b_kmajor=True, c_lower=True, nsplit=4)), (3072, 3072, 3072, dict(a_tri=2, dual=True))])
def test_gemm_tc_3xtf32_matches_fp64(ops, cg, persistent, M, N, K, kw):
    """tcgen05 kind::tf32 x3 on (fp32, lo) operands (the M'^3 products of the backward tail and E E^T) against an fp64 product
    of the same fp32 inputs: 3xTF32 keeps 22 significand bits per operand, so the result is fp32-grade."""
    b_kmajor, a_tri, c_lower = kw.get("b_kmajor", False), kw.get("a_tri", 0), kw.get("c_lower", False)
    alpha, beta, dual, nsplit = kw.get("alpha", 1.0), kw.get("beta", 0.0), kw.get("dual", False), kw.get("nsplit", 1)
    prev = ops.set_tc_cta_group(cg)
    ops.set_tc_persistent(persistent)
    try:
        g = torch.Generator(device="cuda").manual_seed(3 * M + N + K)
        lda, ldn = (K + 7) // 8 * 8, (N + 31) // 32 * 32
        A = torch.randn(M, K, device="cuda", dtype=F32, generator=g)
        A = A.tril() if a_tri == 1 else A.triu() if a_tri == 2 else A
        B = torch.randn((N, K) if b_kmajor else (K, N), device="cuda", dtype=F32, generator=g)
        Af, Bf = _padded(A, lda), _padded(B, lda if b_kmajor else ldn)
        Alo, Blo = _padded(torch.zeros_like(A), lda), _padded(torch.zeros_like(B), lda if b_kmajor else ldn)   # (same leading dimensions)
        ops.split_lo(Af, Alo)
        ops.split_lo(Bf, Blo)
        D, D2 = (_padded(torch.randn(M, N, device="cuda", dtype=F32, generator=g), ldn) for _ in range(2))
        C, C2, Clo, C2lo = (_padded(torch.full((M, N), float("nan"), device="cuda", dtype=F32), ldn) for _ in range(4))
        ref = alpha * (A.double() @ (B.double().T if b_kmajor else B.double())) + beta * D.double()
        ws = torch.empty(nsplit * M * ((N + 3) // 4 * 4), device="cuda") if nsplit > 1 else None
        ops.gemm_tc(Af, Alo, Bf, Blo, C, M, N, K, b_kmajor=b_kmajor, alpha=alpha, beta=beta, D=D if beta else None,
                    C2=C2 if dual else None, D2=D2 if dual else None, a_tri=a_tri, c_lower=c_lower, chunk=2,
                    C_lo=Clo if dual else None, C2_lo=C2lo if dual else None, nsplit=nsplit, split_ws=ws)
        assert (rel(C.tril(), ref.tril()) if c_lower else rel(C, ref)) < 2e-6
        if dual:
            assert rel(C2, ref + D2.double()) < 2e-6
            # the lo companions: what the tensor core does not see of the fp32 outputs (x = trunc_tf32(x) + lo to 2^-22 |x|)
            trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(F32)
            assert float((C - trunc(C) - Clo).abs().max()) <= 2.0 ** -21 * float(C.abs().max())
            assert float((C2 - trunc(C2) - C2lo).abs().max()) <= 2.0 ** -21 * float(C2.abs().max())
    finally:
        ops.set_tc_cta_group(prev)
        ops.set_tc_persistent(1)


def test_tc_operand_scales_and_splits(ops):
    """absmax (order-independent), power-of-two scales from the a-priori bounds, and the two-half split kernels."""
    F16 = torch.float16
    g = torch.Generator().manual_seed(5)
    n = 150
    Ls = (torch.eye(n) + 0.02 * torch.randn(n, n, generator=g)).cuda()
    bits = torch.zeros(4, dtype=torch.int32, device="cuda")
    ops.absmax(Ls, bits[0:1], mode=2)
    E = Ls.tril() - torch.eye(n, device="cuda")
    assert float(bits[0:1].view(torch.float32)) == float(E.abs().max())
    v = torch.randn(1000, generator=g).cuda()
    ops.absmax(v, bits[1:2])
    assert float(bits[1:2].view(torch.float32)) == float(v.abs().max())
    hyp = torch.tensor([0.5, 1.7, 0.3, 0.0, 0, 0, 0, 0], dtype=F64, device="cuda")
    sc = torch.zeros(16, device="cuda")
    ops.tc_scales(hyp, 1e-3, bits, n, sc, 0)
    s = sc.cpu()
    for k in (0, 1, 2, 3, 4):
        assert float(torch.log2(s[k])) == round(float(torch.log2(s[k])))                 # exact powers of two
    a = (1.7 * 4.0) ** 0.5
    assert 1.7 * 8.0 * float(s[1]) <= 2.0 ** 15 < 1.7 * 8.0 * 2.2 * float(s[1])          # K bound os*2/ell^2 -> (2^14, 2^15]
    assert a * float(s[3]) <= 2.0 ** 15 and 1e-3 ** -0.5 * float(s[0]) <= 2.0 ** 15
    assert float(E.abs().max()) * float(s[2]) <= 2.0 ** 15
    assert abs(float(s[8]) * float(s[0]) * float(s[1]) - 1.0) < 1e-6
    # split of tril(Ls) - I and of its transpose
    mk = lambda: torch.zeros(n, 152, dtype=F16, device="cuda")[:, :n]
    hi, lo, hiT, loT = mk(), mk(), mk(), mk()
    ops.split_half(Ls, sc[2:3], hi, lo, mode=2, hiT=hiT, loT=loT)
    back = (hi.double() + lo.double()) / float(s[2])
    assert rel(back, E) < 2.0 ** -21 and float(back.triu(1).abs().max()) == 0
    assert torch.equal(hiT, hi.T) and torch.equal(loT, lo.T)
    W = torch.randn(70, 70, generator=g, dtype=F64).cuda()
    mk2 = lambda: torch.zeros(70, 72, dtype=F16, device="cuda")[:, :70]
    wh, wl = mk2(), mk2()
    ops.split_half(W, sc[0:1], wh, wl, mode=1)
    assert rel((wh.double() + wl.double()) / float(s[0]), W.tril()) < 2.0 ** -21


@pytest.mark.parametrize("dtype", [F32, F64])
@pytest.mark.parametrize("n,pad", [(1, 0), (7, 1), (96, 0), (385, 3), (1024, 0)])
def test_phi_outer_matches_the_three_pass_form(ops, dtype, n, pad):
    """lower triangle of Phi(X + u v^T) in fp64 (one pass) = add_outer -> cast -> phi_lower; the upper triangle is not touched"""
    g = torch.Generator(device="cuda").manual_seed(n)
    X = torch.randn(n, n + pad, device="cuda", dtype=dtype, generator=g)[:, :n]
    u, v = (torch.randn(n, device="cuda", dtype=dtype, generator=g) for _ in range(2))
    P = torch.full((n, n + pad), float("nan"), dtype=F64, device="cuda")[:, :n]
    ops.phi_outer(X, u, v, P, n)
    ref = (X + torch.outer(u, v)).double().tril()
    ref.diagonal().mul_(0.5)
    assert bool(torch.isnan(P.triu(1)[torch.ones(n, n, dtype=torch.bool, device="cuda").triu(2)]).all()) or n < 3
    low = torch.ones(n, n, dtype=torch.bool, device="cuda").tril()
    t = 1e-14 if dtype == F64 else 1e-6
    assert float((P[low] - ref[low]).abs().max()) <= t * max(1.0, float(ref.abs().max()))
