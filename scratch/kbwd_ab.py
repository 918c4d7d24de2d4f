"""A/B of the two thread mappings of kdir_bwd_v4 (4 or 2 column points per lane) at C3 / C5 shapes: time and agreement."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (M, n, d, p1, p2) in ((1024, 16384, 10, 2, 2), (1024, 16384, 18, 2, 0), (512, 4096, 3, 1, 1)):
    Z = torch.randn(M, d, device=dev); x = torch.randn(n, d, device=dev)
    Vz = torch.randn(M * p1, d, device=dev); Vx = torch.randn(max(n * p2, 1), d, device=dev)
    hyp = torch.tensor([1.3, 0.9, 0.1, 0.0], dtype=torch.float64, device=dev)
    u, inv = ops.normalize_dirs(Vz)
    w, _ = ops.normalize_dirs(Vx)
    ld = ((n * (p2 + 1) + 63) // 64) * 64
    dK = torch.randn(M * (p1 + 1), ld, device=dev)
    outs = {}
    for vpl in (4, 2):
        ops.set_kdir_bwd_vpl(vpl)
        gx = torch.zeros(M, d, dtype=torch.float64, device=dev); gv = torch.zeros(M * p1, d, dtype=torch.float64, device=dev)
        gsc = torch.zeros(2, dtype=torch.float64, device=dev)
        ops.kdir_bwd(Z, u, inv, p1, x, w, p2, hyp, dK[:, : n * (p2 + 1)], gx, gv, gsc)
        torch.cuda.synchronize()
        outs[vpl] = (gx.clone(), gv.clone(), gsc.clone())
        ms = t(lambda: ops.kdir_bwd(Z, u, inv, p1, x, w, p2, hyp, dK[:, : n * (p2 + 1)], gx, gv, gsc))
        print(f"M={M} n={n} d={d} p=({p1},{p2}) vpl={vpl}: {ms:.3f} ms  ({dK.numel()*4/ms/1e6:.0f} GB/s of dK)", flush=True)
    rel = lambda a, b: float((a - b).abs().max() / a.abs().max())
    print("   agreement 4 vs 2:", [rel(a, b) for a, b in zip(outs[4], outs[2])])
ops.set_kdir_bwd_vpl(2)
