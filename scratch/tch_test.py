"""GPU scratch: correctness + speed of the tcgen05 3xFP16 GEMM against fp64 torch and the 3xTF32 kernel."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch
from dsvgp_b200 import ops
F16, F32, F64 = torch.float16, torch.float32, torch.float64
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
torch.manual_seed(0)
CG = int(sys.argv[1]) if len(sys.argv) > 1 else 2
print('cta_group', ops.set_tc_cta_group(CG))
def split(x, s):
    xs = x.float() * s
    hi = xs.half(); lo = (xs - hi.float()).half()
    return hi, lo
def pad(t, ld):
    out = torch.zeros(t.shape[0], ld, device=t.device, dtype=t.dtype); out[:, :t.shape[1]] = t; return out
def run(M, N, K, b_kmajor=False, a_tri=0, c_lower=False, chunk=1, alpha=1.0, beta=0.0, dual=False, nsplit=1):
    A = torch.randn(M, K, device="cuda", dtype=F64)
    if a_tri == 1: A = A.tril()
    if a_tri == 2: A = A.triu()
    B = torch.randn(N, K, device="cuda", dtype=F64) if b_kmajor else torch.randn(K, N, device="cuda", dtype=F64)
    sA, sB = 2.0 ** 11, 2.0 ** 12
    lda = (K + 7) // 8 * 8
    ldb = lda if b_kmajor else (N + 63) // 64 * 64
    Ah, Al = (pad(t, lda)[:, :K] for t in split(A, sA))
    Bh, Bl = (pad(t, ldb)[:, :B.shape[1]] for t in split(B, sB))
    ldc = (N + 63) // 64 * 64
    D = torch.randn(M, ldc, device="cuda", dtype=F32)[:, :N]; D2 = torch.randn(M, ldc, device="cuda", dtype=F32)[:, :N]
    C = torch.full((M, ldc), float("nan"), device="cuda", dtype=F32)[:, :N]
    C2 = torch.full((M, ldc), float("nan"), device="cuda", dtype=F32)[:, :N]
    Ch = tuple(torch.zeros(M, ldc, device="cuda", dtype=F16)[:, :N] for _ in range(2))
    C2h = tuple(torch.zeros(M, ldc, device="cuda", dtype=F16)[:, :N] for _ in range(2))
    inv = torch.tensor([1.0 / (sA * sB)], device="cuda", dtype=F32)
    cs, c2s = torch.tensor([2.0 ** 3], device="cuda"), torch.tensor([2.0 ** 2], device="cuda")
    Aq = (Ah.double() + Al.double()) / sA; Bq = (Bh.double() + Bl.double()) / sB     # what the kernel is given
    ref = alpha * (Aq @ (Bq.T if b_kmajor else Bq)) + beta * D.double()
    ws = torch.empty(nsplit * M * ((N + 3) // 4 * 4), device="cuda") if nsplit > 1 else None
    ops.gemm_tch((Ah, Al), (Bh, Bl), C, M, N, K, inv, b_kmajor=b_kmajor, alpha=alpha, beta=beta, D=D if beta else None,
                 C2=C2 if dual else None, D2=D2 if dual else None, Ch=Ch if (dual and nsplit == 1) else None, c_scale=cs,
                 C2h=C2h if dual else None, c2_scale=c2s, a_tri=a_tri, c_lower=c_lower, chunk=chunk, nsplit=nsplit, split_ws=ws)
    torch.cuda.synchronize()
    e = rel(C.tril(), ref.tril()) if c_lower else rel(C, ref)
    e2 = eh = 0.0
    if dual:
        e2 = rel(C2, ref + D2.double())
        eh = max(rel((Ch[0].double() + Ch[1].double()) / 8.0, ref), rel((C2h[0].double() + C2h[1].double()) / 4.0, ref + D2.double()))
    print(f"M{M} N{N} K{K} kmaj{int(b_kmajor)} tri{a_tri} clow{int(c_lower)} chunk{chunk} a{alpha} b{beta} split{nsplit}: err {e:.2e} dual {e2:.2e} half-out {eh:.2e}", flush=True)
run(128, 256, 64)
run(128, 256, 128)
run(128, 256, 256, chunk=2)
run(256, 512, 512)
run(200, 300, 100)
run(384, 1000, 384, a_tri=1)
run(384, 1000, 384, a_tri=2)
run(384, 1000, 384, a_tri=1, alpha=2.0, beta=-2.0, dual=True)
run(300, 300, 1000, b_kmajor=True)
run(512, 512, 2048, b_kmajor=True, c_lower=True)
run(512, 512, 8192, b_kmajor=True, c_lower=True, nsplit=4)
run(3072, 4096, 3072, a_tri=1)
# speed at the bench shape
M = K = 3072; N = 49152
A = torch.randn(M, K, device="cuda").tril(); B = torch.randn(K, N, device="cuda")
Ah, Al = split(A, 2.0 ** 10); Bh, Bl = split(B, 2.0 ** 10)
inv = torch.tensor([2.0 ** -20], device="cuda"); C = torch.empty(M, N, device="cuda")
Alo32, Blo32 = ops.split_lo(A), ops.split_lo(B)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for chunk in (1, 2, 4):
    ms = t(lambda: ops.gemm_tch((Ah, Al), (Bh, Bl), C, M, N, K, inv, a_tri=1, chunk=chunk))
    print(f"3xFP16 trmm chunk {chunk}: {ms:.3f} ms  {M*K*N/ms/1e9:.1f} useful TFLOP/s", flush=True)
ms = t(lambda: ops.gemm_tc(A, Alo32, B, Blo32, C, M, N, K, a_tri=1, chunk=2))
print(f"3xTF32 trmm chunk 2: {ms:.3f} ms  {M*K*N/ms/1e9:.1f} useful TFLOP/s", flush=True)
ref = (A.double() @ B[:, :2048].double())
ops.gemm_tch((Ah, Al), (Bh, Bl), C, M, N, K, inv, a_tri=1, chunk=1); print("big err 3xFP16 chunk1 vs exact fp32 inputs", rel(C[:, :2048], ref))
ops.gemm_tc(A, Alo32, B, Blo32, C, M, N, K, a_tri=1, chunk=2); print("big err 3xTF32 chunk2", rel(C[:, :2048], ref))
