"""kdir_fwd_v4: row points per CTA (TIB) against the grid size -- small minibatches (C4: n = 2048) leave the GPU half empty at 64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
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
for name, n in (("C4", 2048), ("C4", 512), ("C3", 512), ("C3", 4096), ("C3", 16384), ("C5", 512), ("C5", 16384)):
    wl = bench.WORKLOADS[name]
    M, d, p = wl["M"], wl["d"], wl["p"]
    p2 = 0 if wl["variant"] == "dfree" else p
    Z, x = torch.rand(M, d, device=dev), torch.rand(n, d, device=dev)
    Vz = torch.randn(M * p, d, device=dev)
    idx = torch.stack([torch.randperm(d)[:max(p2, 1)] for _ in range(n)]).reshape(-1)
    Vx = torch.eye(d, device=dev)[idx.to(dev)]
    u = ops.normalize_dirs(Vz)[0]
    hyp = torch.tensor([0.7, 0.9, 0.1, 0.0, 0.5, 0.5, 0.5, 0.0], dtype=torch.float64, device=dev)
    if p2:
        w, _, cidx, flag = ops.normalize_dirs_canon(Vx); canon = (cidx, flag)
    else:
        w, canon = None, None
    ld = ((n * (p2 + 1) + 63) // 64) * 64
    K = torch.empty(M * (p + 1), ld, device=dev)[:, : n * (p2 + 1)]
    res = []
    for tib in (64, 32, 16, 8):
        ops.set_kdir_fwd_knobs(tib, 2)
        res.append((tib, t(lambda: ops.kdir_fwd(Z, u, p, x, w, p2, hyp, K, canon=canon))))
    ops.set_kdir_fwd_knobs(0, 2)      # back to the adaptive default
    print(f"{name} n={n}: " + "  ".join(f"TIB {a}: {b:6.1f} us" for a, b in res), flush=True)
