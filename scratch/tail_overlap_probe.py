"""Feasibility probe: the fp64 Cholesky-backward products (DMMA pipe) beside the K_zx assembly backward (FP32 FMA issue) on two
streams -- do they overlap on the same SMs, or just split the GPU?  Makespan of both against the sum of each alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from dsvgp_b200 import ops, engine
from dsvgp_b200.engine import ENGINE
dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS["C3"])
arm = bench.Arm(wl, dev, 0, 1)
x, V, y = (t.to(dev) for t in arm.batch(wl["n"], 1))
for _ in range(3): arm.step(x, V, y)
torch.cuda.synchronize()
ws = ENGINE.workspace(dev, torch.float32, wl["n"], wl["d"], wl["M"], wl["p"], wl["p"])
f = ENGINE.factor(dev, torch.float32, wl["d"], wl["M"], wl["p"])
Mq, nq, p = ws.Mq, ws.nq, wl["p"]
Z = arm.model.variational_strategy.inducing_points.detach()
wx = ENGINE._data_dirs(ws, V, torch.float32)
gZ, gV, gs = torch.zeros_like(ws.gZ), torch.zeros_like(ws.gVz), torch.zeros(2, dtype=torch.float64, device=dev)
def kbwd():
    ops.kdir_bwd(Z, f.uzT, f.invzT, p, x, wx, p, f.hyp, ws.Kzx, gZ, gV, gs)
W = f.W
def tail():
    ops.gemm(ws.Psi, W, ws.Y, a_tri=ops.TRI_LOWER, b_tri=ops.TRI_LOWER, c_tri=1, alpha=-1.0, M=Mq, N=Mq, K=Mq)
    ops.gemm(W, ws.Y, ws.S, ta=True, a_tri=ops.TRI_UPPER, b_tri=ops.TRI_LOWER, M=Mq, N=Mq, K=Mq)
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
def timed(fa, fb):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    s1.wait_stream(cur); s2.wait_stream(cur)
    if fa:
        with torch.cuda.stream(s1): fa()
    if fb:
        with torch.cuda.stream(s2): fb()
    cur.wait_stream(s1); cur.wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
med = lambda fa, fb: sorted(timed(fa, fb) for _ in range(7))[3]
for vpl in (4, 2):
    ops.set_kdir_bwd_vpl(vpl)
    kbwd(); tail(); torch.cuda.synchronize()
    a, b, ab = med(kbwd, None), med(None, tail), med(kbwd, tail)
    print(f"vpl={vpl}: assembly backward alone {a:.3f} ms, fp64 tail products alone {b:.3f} ms, both on two streams {ab:.3f} ms (sum {a + b:.3f})", flush=True)
ops.set_kdir_bwd_vpl(4)
