"""Feasibility probe: how much does a tensor-core product that runs BESIDE the second half of the factorisation (on a limited
number of CTA pairs) slow the factorisation down, and how long does it take itself?"""
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
Mq, nq, sc, H = ws.Mq, ws.nq, f.scales, engine.TCH_CHUNK
Z = arm.model.variational_strategy.inducing_points.detach()
def chol():
    ops.kdir_fwd(Z, f.uz64, wl["p"], Z, f.uz64, wl["p"], f.hyp, f.Kzz, diag_add=1e-3)
    ops.cholesky_inverse(f.Kzz, f.L, f.W, f.nb0, f.nlev, f.info)
half = Mq // 2
def a_top():
    ops.gemm_tch((f.Wh, f.Wl), (ws.Kh, ws.Kl), ws.A, half, nq, half, sc[8:9], a_tri=ops.TRI_LOWER, chunk=H, Ch=(ws.Ah, ws.Al), c_scale=sc[3:4])
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
def run(pairs, delay_ms):
    ops.set_tc_max_pairs(pairs if pairs else 0)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(s1):
        e[0].record(); chol(); e[1].record()
    if pairs is not None:
        with torch.cuda.stream(s2):
            torch.cuda._sleep(int(delay_ms * 1.9e6))
            e[2].record(); a_top(); e[3].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), (e[2].elapsed_time(e[3]) if pairs is not None else 0.0)
for _ in range(2): run(None, 0); run(0, 1.3)
for pairs in (None, 0, 64, 56, 48, 40):
    r = [run(pairs, 1.3) for _ in range(5)]
    c = sorted(v[0] for v in r)[2]; g = sorted(v[1] for v in r)[2]
    print(f"A_top on {'-' if pairs is None else (pairs or 74)} pairs beside the factorisation: factorisation {c:.3f} ms, A_top {g:.3f} ms", flush=True)
ops.set_tc_max_pairs(0)
