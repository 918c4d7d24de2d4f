"""Assembly cost versus the number of directions p = d of the full-gradient strategy's kernel (blocked kernels for p <= 3, runtime-p generic beyond)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops
dev = torch.device("cuda", 0)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
M, n = 256, 2048
for d in (2, 3, 4, 5, 6, 8, 10):
    p = d
    Z = torch.rand(M, d, device=dev); x = torch.rand(n, d, device=dev)
    Vz = torch.eye(d, device=dev).repeat(M, 1); Vx = torch.eye(d, device=dev).repeat(n, 1)
    hyp = torch.tensor([0.7, 0.9, 0.1, 0.0, 0.5, 0.5, 0.5, 0.0], dtype=torch.float64, device=dev)
    u, inv = ops.normalize_dirs(Vz); w, _ = ops.normalize_dirs(Vx)
    K = torch.empty(M * (p + 1), ((n * (p + 1) + 63) // 64) * 64, device=dev)[:, : n * (p + 1)]
    ms_f = t(lambda: ops.kdir_fwd(Z, u, p, x, w, p, hyp, K))
    dK = torch.randn_like(K)
    gx = torch.zeros(M, d, dtype=torch.float64, device=dev); gv = torch.zeros(M * p, d, dtype=torch.float64, device=dev); gsc = torch.zeros(2, dtype=torch.float64, device=dev)
    ms_b = t(lambda: ops.kdir_bwd(Z, u, inv, p, x, w, p, hyp, dK, gx, gv, gsc))
    by = K.numel() * 4
    print(f"d=p={d}: K {M*(p+1)} x {n*(p+1)}  fwd {ms_f:.3f} ms ({by/ms_f/1e6:.0f} GB/s)  bwd {ms_b:.3f} ms ({by/ms_b/1e6:.0f} GB/s)", flush=True)
