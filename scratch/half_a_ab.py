"""A/B of engine.HALF_A (whitening product writes only the two-half split of A) on the C3 step, plus the two passes that read A."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from dsvgp_b200 import ops, engine
from dsvgp_b200.engine import ENGINE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
wl = dict(bench.WORKLOADS["C3"], n=n)
dev = torch.device("cuda", 0)
arm = bench.Arm(wl, dev, 0, 1)
x, V, y = (t.to(dev) for t in arm.batch(n, 1))
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for rnd in range(2):
    for flag in (False, True):
        engine.HALF_A = flag
        print(f"HALF_A={flag}: step {t(lambda: arm.step(x, V, y)):.3f} ms", flush=True)
ws = ENGINE.workspace(dev, torch.float32, n, wl["d"], wl["M"], wl["p"], wl["p"])
f = ENGINE.factor(dev, torch.float32, wl["d"], wl["M"], wl["p"])
Mq, nq, sc = ws.Mq, ws.nq, f.scales
m = torch.randn(Mq, device=dev); gmu = torch.randn(nq, device=dev); gvar = torch.randn(nq, device=dev)
ws.C.normal_(); ws.A.normal_()
one = torch.full((1,), 1024.0, device=dev)
ops.split_half(ws.A, one, ws.Ah, ws.Al, rows=Mq, cols=nq)
cm = torch.zeros(1, dtype=torch.int32, device=dev)
print("col_dots fp32 A     ", t(lambda: ops.col_dots(ws.A, m, ws.pm, ws.pv, Mq, nq, C=ws.C, cmax=cm)))
pm0, pv0 = ws.pm.clone(), ws.pv.clone()
print("col_dots half A     ", t(lambda: ops.col_dots_half((ws.Ah, ws.Al), one, m, ws.pm, ws.pv, Mq, nq, ws.C, cmax=cm)))
print("   rel diff pm, pv:", float((ws.pm - pm0).abs().max() / pm0.abs().max()), float((ws.pv - pv0).abs().max() / pv0.abs().max()))
print("dA_half fp32 A      ", t(lambda: ops.dA_apply_half(ws.A, ws.C, Mq, nq, m, gmu, gvar, ws.tp, ws.t, ws.Kh, ws.Kl, ws.Agh, ws.Agl, sc[5:6], sc[6:7])))
print("dA_half half A      ", t(lambda: ops.dA_apply_half(None, ws.C, Mq, nq, m, gmu, gvar, ws.tp, ws.t, ws.Kh, ws.Kl, ws.Agh, ws.Agl, sc[5:6], sc[6:7], A_half=(ws.Ah, ws.Al), a_scale=one)))
H = engine.TCH_CHUNK
print("A product fp32+split", t(lambda: ops.gemm_tch((f.Wh, f.Wl), (ws.Kh, ws.Kl), ws.A, Mq, nq, Mq, sc[8:9], a_tri=ops.TRI_LOWER, chunk=H, Ch=(ws.Ah, ws.Al), c_scale=sc[3:4])))
print("A product split only", t(lambda: ops.gemm_tch((f.Wh, f.Wl), (ws.Kh, ws.Kl), None, Mq, nq, Mq, sc[8:9], a_tri=ops.TRI_LOWER, chunk=H, Ch=(ws.Ah, ws.Al), c_scale=sc[3:4])))
