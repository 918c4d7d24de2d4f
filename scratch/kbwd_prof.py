"""One C3-shaped call of the vectorised assembly backward, for ncu:
   ncu --set full --import-source on --clock-control none -k regex:kdir_bwd_v4 --launch-skip 2 -c 1 -o gpurun_out/kbwd python scratch/kbwd_prof.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
M, n, d, p1, p2 = 1024, 16384, 10, 2, 2
Z = torch.randn(M, d, device=dev); x = torch.randn(n, d, device=dev)
Vz = torch.randn(M * p1, d, device=dev); Vx = torch.randn(n * p2, d, device=dev)
hyp = torch.tensor([1.3, 0.9, 0.1, 0.0], dtype=torch.float64, device=dev)
u, inv = ops.normalize_dirs(Vz)
w, _ = ops.normalize_dirs(Vx)
ld = ((n * (p2 + 1) + 63) // 64) * 64
dK = torch.randn(M * (p1 + 1), ld, device=dev)
gx = torch.zeros(M, d, dtype=torch.float64, device=dev); gv = torch.zeros(M * p1, d, dtype=torch.float64, device=dev)
gsc = torch.zeros(2, dtype=torch.float64, device=dev)
for _ in range(4):
    ops.kdir_bwd(Z, u, inv, p1, x, w, p2, hyp, dK[:, : n * (p2 + 1)], gx, gv, gsc)
torch.cuda.synchronize()
