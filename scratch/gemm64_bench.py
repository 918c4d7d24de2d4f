"""fp64 DMMA GEMM (csrc/gemm.cu) on the shapes the step uses: TFLOP/s against the measured DMMA peak."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops
F64 = torch.float64
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(n, n, dtype=F64, device="cuda", generator=g)
B = torch.randn(n, n, dtype=F64, device="cuda", generator=g)
C = torch.zeros(n, n, dtype=F64, device="cuda")
P = torch.randn(n, 96, dtype=F64, device="cuda", generator=g)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cases = {
    "dense NN": (lambda: ops.gemm(A, B, C), 2.0 * n ** 3),
    "Phi W (lower x lower -> lower)": (lambda: ops.gemm(A, B, C, a_tri=ops.TRI_LOWER, b_tri=ops.TRI_LOWER, c_tri=1, alpha=-1.0), 2.0 * n ** 3 / 6),
    "W^T Y (upper^T x lower)": (lambda: ops.gemm(A, B, C, ta=True, a_tri=ops.TRI_UPPER, b_tri=ops.TRI_LOWER), 2.0 * n ** 3 / 3),
    "W K (lower x dense n x 1536)": (lambda: ops.gemm(A, B[:, :1536], C[:, :1536], a_tri=ops.TRI_LOWER, N=1536), 1.0 * n * n * 1536),
    "trailing update (n x n x 96, lower)": (lambda: ops.gemm(P, P, C, tb=True, alpha=-1.0, beta=1.0, c_tri=1), 1.0 * n * n * 96),
    "inverse top level (lower-tri A, n/2)": (lambda: ops.gemm(A[: n // 2, : n // 2], B[: n // 2, : n // 2], C[: n // 2, : n // 2], a_tri=ops.TRI_LOWER, alpha=-1.0), 1.0 * (n // 2) ** 3),
}
ref = torch.matmul(A.tril(), B.tril()).tril()
ops.gemm(A, B, C, a_tri=ops.TRI_LOWER, b_tri=ops.TRI_LOWER, c_tri=1)
print("check lower x lower:", float((C.tril() - ref).abs().max() / ref.abs().max()))
for name, (fn, flops) in cases.items():
    ms = t(fn)
    print(f"{name:40s} {ms:7.3f} ms  {flops / ms / 1e9:6.1f} TFLOP/s")
from dsvgp_b200 import _lib
print("--- register-staged kernel of round 1 (dsvgp_set_gemm64_async(0))")
_lib.call_raw("dsvgp_set_gemm64_async", 0)
for name, (fn, flops) in cases.items():
    ms = t(fn)
    print(f"{name:40s} {ms:7.3f} ms  {flops / ms / 1e9:6.1f} TFLOP/s")
_lib.call_raw("dsvgp_set_gemm64_async", 1)
