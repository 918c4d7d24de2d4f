"""Small calls of the vectorised assembly kernels (forward, backward; whole and ragged column tiles) for compute-sanitizer:
   compute-sanitizer --tool memcheck|racecheck python scratch/sanitize_kdir.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
hyp = torch.tensor([1.3, 0.9, 0.1, 0.0, 0.5, 0.5, 0.5, 0.0], dtype=torch.float64, device=dev)
for (M, n, d, p1, p2) in ((70, 1024, 10, 2, 2), (33, 1000, 10, 2, 2), (65, 640, 18, 2, 0), (40, 777, 3, 1, 1)):
    Z = torch.randn(M, d, device=dev); x = torch.randn(n, d, device=dev)
    Vz = torch.randn(M * p1, d, device=dev); Vx = torch.randn(max(n * p2, 1), d, device=dev)
    u, inv = ops.normalize_dirs(Vz)
    w = ops.normalize_dirs(Vx)[0] if p2 else None
    ld = ((n * (p2 + 1) + 63) // 64) * 64
    K = torch.empty(M * (p1 + 1), ld, device=dev)[:, : n * (p2 + 1)]
    ops.kdir_fwd(Z, u, p1, x, w, p2, hyp, K)
    dK = torch.randn(M * (p1 + 1), ld, device=dev)[:, : n * (p2 + 1)]
    gx = torch.zeros(M, d, dtype=torch.float64, device=dev); gv = torch.zeros(M * p1, d, dtype=torch.float64, device=dev)
    gsc = torch.zeros(2, dtype=torch.float64, device=dev)
    for vpl in (4, 2):
        ops.set_kdir_bwd_vpl(vpl)
        ops.kdir_bwd(Z, u, inv, p1, x, w, p2, hyp, dK, gx, gv, gsc)
    ops.set_kdir_bwd_vpl(4)
    torch.cuda.synchronize()
    print("ok", M, n, d, p1, p2, float(gx.abs().sum()), flush=True)
