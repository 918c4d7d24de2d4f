"""K_zx assembly of the C2 / C4 / C5 shapes, one launch each (for an ncu capture: the bound each of them sits under)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from dsvgp_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
cases = []
for name in ("C2", "C4", "C5"):
    wl = bench.WORKLOADS[name]
    T = torch.float64 if wl["dtype"] == "f64" else torch.float32
    M, n, d, p = wl["M"], wl["n"], wl["d"], wl["p"]
    p2 = 0 if wl["variant"] == "dfree" else p
    Z, x = torch.rand(M, d, device=dev, dtype=T), torch.rand(n, d, device=dev, dtype=T)
    Vz = torch.randn(M * p, d, device=dev, dtype=T)
    idx = torch.stack([torch.randperm(d)[:max(p2, 1)] for _ in range(n)]).reshape(-1)
    Vx = torch.eye(d, device=dev, dtype=T)[idx.to(dev)]
    u = ops.normalize_dirs(Vz)[0]
    hyp = torch.tensor([0.7, 0.9, 0.1, 0.0, 0.5, 0.5, 0.5, 0.0], dtype=torch.float64, device=dev)
    if T == torch.float32 and p2:
        w, _, cidx, flag = ops.normalize_dirs_canon(Vx); canon = (cidx, flag)
    else:
        w, canon = (ops.normalize_dirs(Vx)[0] if p2 else None), None
    ld = ((n * (p2 + 1) + 63) // 64) * 64
    K = torch.empty(M * (p + 1), ld, device=dev, dtype=T)[:, : n * (p2 + 1)]
    cases.append((Z, u, p, x, w, p2, hyp, K, canon))
run = lambda c: ops.kdir_fwd(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], canon=c[8])
for c in cases: run(c); run(c)
torch.cuda.synchronize(); torch.cuda.profiler.start()
for c in cases: run(c)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
