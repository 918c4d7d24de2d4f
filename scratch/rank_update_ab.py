"""The fp64 rank-96 trailing update of the blocked Cholesky alone: general kernel vs rank-update kernel, at the sizes of links 0, 8, 16, 24."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops
dev = torch.device("cuda", 0)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
nb = 96
L = torch.randn(3072, 3072, dtype=torch.float64, device=dev)
C0 = torch.randn(3072, 3072, dtype=torch.float64, device=dev)
for k in (0, 8, 16, 24):
    m = 3072 - (k + 1) * nb; m2 = m - nb
    A, B = L[(k + 2) * nb:, k * nb:(k + 1) * nb], L[(k + 1) * nb:, k * nb:(k + 1) * nb]
    for ru in (0, 1):
        ops.set_rank_update(ru)
        C = C0.clone()
        Cv = C[(k + 2) * nb:, (k + 1) * nb:]
        us = t(lambda: ops.gemm(A, B, Cv, tb=True, alpha=-1.0, beta=1.0, c_tri=1, M=m2, N=m, K=nb))
        fl = 2.0 * nb * m2 * m / 2
        print(f"link {k}: m2={m2} rank_update={ru}: {us:7.1f} us  ({fl / us / 1e6:5.1f} TFLOP/s on the lower half)", flush=True)
ops.set_rank_update(0)
