"""clock64() stamps at the phase boundaries of the variant-2 diagonal-block kernel (last block of a 3072 factorisation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops, _lib
Mq = 3072
F64 = torch.float64
g = torch.Generator().manual_seed(Mq)
R = torch.randn(Mq, Mq + 5, generator=g, dtype=F64)
A = (R @ R.T / (Mq + 5) + 1e-3 * torch.eye(Mq, dtype=F64)).cuda()
Mp, nb0, nlev = ops.chol_plan(Mq)
Aw = torch.empty(Mp, Mp, dtype=F64, device="cuda"); L = torch.empty_like(Aw); W = torch.empty_like(Aw)
info = torch.ones(1, dtype=torch.int32, device="cuda")
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.call_raw("dsvgp_set_potrf_debug", dbg)
for _ in range(3):
    Aw.zero_(); Aw[:Mq, :Mq] = A; ops.pad_identity(Aw, Mq)
    ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
torch.cuda.synchronize()
t = dbg.cpu().tolist()
print("cluster kernel: entry->loads", t[52]-t[61], "Z", t[53]-t[52], "SYRK", t[54]-t[53], "barrier1", t[55]-t[54], "gather", t[56]-t[55], "barrier2", t[57]-t[56], "zero Ws + first sync", t[1]-t[57])
for j in range(12):
    s1w0 = t[2 + 4 * j] - (t[1] if j == 0 else t[5 + 4 * (j - 1)])
    s1 = t[3 + 4 * j] - (t[1] if j == 0 else t[5 + 4 * (j - 1)])
    s2 = (t[4 + 4 * j] - t[3 + 4 * j]) if j < 11 else 0
    s3 = (t[5 + 4 * j] - t[4 + 4 * j]) if j < 11 else 0
    print(f"step {j:2d}: S1 warp0 {s1w0:6d}  S1 all {s1:6d}  S2 {s2:5d}  S3 {s3:5d}")
print("last inverse row", t[62] - t[3 + 44], " write-out", t[63] - t[62], " total", t[63] - t[61])
_lib.call_raw("dsvgp_set_potrf_debug", None)
