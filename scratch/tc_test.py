"""GPU scratch: correctness + speed of the tcgen05 3xTF32 GEMM against fp64 torch and the mma.sync kernel."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch
from dsvgp_b200 import ops
F32, F64 = torch.float32, torch.float64
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
torch.manual_seed(0)
CG = int(sys.argv[1]) if len(sys.argv) > 1 else 1
print('cta_group', ops.set_tc_cta_group(CG))
def run(M, N, K, b_kmajor=False, a_tri=0, c_lower=False, chunk=2, alpha=1.0, beta=0.0, dual=False, ldpad=0):
    A = torch.randn(M, K, device="cuda", dtype=F64)
    if a_tri == 1: A = A.tril()
    if a_tri == 2: A = A.triu()
    ldb = ((N + 31) // 32) * 32 + ldpad
    if b_kmajor:
        Bfull = torch.randn(N, ((K + 7) // 8) * 8, device="cuda", dtype=F64); B = Bfull[:, :K]; Bm = B.T
    else:
        Bfull = torch.randn(K, ldb, device="cuda", dtype=F64); B = Bfull[:, :N]; Bm = B
    lda = ((K + 7) // 8) * 8
    Af = torch.zeros(M, lda, device="cuda", dtype=F32); Af[:, :K] = A.float(); Av = Af[:, :K]
    Bf = Bfull.float(); Bv = Bf[:, :K] if b_kmajor else Bf[:, :N]
    Alo, Blo = torch.empty_like(Af), torch.empty_like(Bf)
    ops.split_lo(Af, Alo); ops.split_lo(Bf, Blo)
    D = torch.randn(M, ldb if not b_kmajor else N, device="cuda", dtype=F32)
    D2 = torch.randn_like(D)
    C = torch.full_like(D, float("nan")); C2 = torch.full_like(D, float("nan"))
    ref = alpha * (Av.double() @ (Bv.double().T if b_kmajor else Bv.double())) + beta * D[:, :N].double()
    ops.gemm_tc(Af, Alo, Bf, Blo, C, M, N, K, b_kmajor=b_kmajor, alpha=alpha, beta=beta, D=D if beta else None,
                C2=C2 if dual else None, D2=D2 if dual else None, a_tri=a_tri, c_lower=c_lower, chunk=chunk)
    torch.cuda.synchronize()
    out = C[:, :N]
    if c_lower:
        e = rel(out.tril(), ref.tril())
    else:
        e = rel(out, ref)
    e2 = rel(C2[:, :N], ref + D2[:, :N].double()) if dual else 0.0
    print(f"M{M} N{N} K{K} kmaj{int(b_kmajor)} tri{a_tri} clow{int(c_lower)} chunk{chunk} a{alpha} b{beta}: err {e:.2e} dual {e2:.2e}", flush=True)
    return e
run(128, 256, 32)
run(128, 256, 64)
run(128, 256, 256)
run(256, 512, 512)
run(200, 300, 100)
run(128, 256, 128, chunk=1)
run(384, 1000, 384, a_tri=1)
run(384, 1000, 384, a_tri=2)
run(384, 1000, 384, a_tri=1, alpha=2.0, beta=-2.0, dual=True)
run(300, 300, 1000, b_kmajor=True)
run(512, 512, 2048, b_kmajor=True, c_lower=True)
run(3072, 4096, 3072, a_tri=1)
# speed at the bench shape
M = K = 3072; N = 49152
A = torch.randn(M, K, device="cuda").tril(); B = torch.randn(K, N, device="cuda")
Alo, Blo = ops.split_lo(A), ops.split_lo(B); C = torch.empty(M, N, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for chunk in (1, 2, 4, 8, 1000):
    ms = t(lambda: ops.gemm_tc(A, Alo, B, Blo, C, M, N, K, a_tri=1, chunk=chunk))
    print(f"tc trmm chunk {chunk}: {ms:.3f} ms  {M*K*N/ms/1e9:.1f} useful TFLOP/s", flush=True)
ms = t(lambda: ops.gemm(A, B, C, a_tri=ops.TRI_LOWER))
print(f"mma.sync trmm: {ms:.3f} ms  {M*K*N/ms/1e9:.1f} useful TFLOP/s")
ms = t(lambda: ops.split_lo(B, Blo)); print(f"split_lo big: {ms:.3f} ms  {3*B.numel()*4/ms/1e6:.0f} GB/s")
ref = (A.double() @ B[:, :2048].double())
ops.gemm_tc(A, Alo, B, Blo, C, M, N, K, a_tri=1, chunk=2); print("big err chunk2", rel(C[:, :2048], ref))
ops.gemm_tc(A, Alo, B, Blo, C, M, N, K, a_tri=1, chunk=1000); print("big err chunk inf", rel(C[:, :2048], ref))
ops.gemm(A, B, C, a_tri=ops.TRI_LOWER); print("big err mma.sync+fp64 master", rel(C[:, :2048], ref))
