"""A/B of the persistent (one CTA pair per SM pair, balanced work list) and the per-tile CTA-pair kernels of the 3xFP16
tcgen05 product on the four big products of a C3 step (n = 16384), plus the whole step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from dsvgp_b200 import ops, gp, engine
from dsvgp_b200.engine import ENGINE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
wl = dict(bench.WORKLOADS["C3"], n=n)
dev = torch.device("cuda", 0)
arm = bench.Arm(wl, dev, 0, 1)
x, V, y = (t.to(dev) for t in arm.batch(n, 1))
for _ in range(3): arm.step(x, V, y)
torch.cuda.synchronize()
ws = ENGINE.workspace(dev, torch.float32, n, wl["d"], wl["M"], wl["p"], wl["p"])
f = ENGINE.factor(dev, torch.float32, wl["d"], wl["M"], wl["p"])
Mq, nq, sc, H = ws.Mq, ws.nq, f.scales, engine.TCH_CHUNK
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
prods = {
 "A = W K_zx (lower tri)": lambda: ops.gemm_tch((f.Wh, f.Wl), (ws.Kh, ws.Kl), ws.A, Mq, nq, Mq, sc[8:9], a_tri=ops.TRI_LOWER, chunk=H, Ch=(ws.Ah, ws.Al), c_scale=sc[3:4]),
 "C = D A (dense)": lambda: ops.gemm_tch((ws.Dh, ws.Dl), (ws.Ah, ws.Al), ws.C, Mq, nq, Mq, sc[13:14], chunk=H),
 "dK_zx = W^T dA (upper tri)": lambda: ops.gemm_tch((f.WTh, f.WTl), (ws.Kh, ws.Kl), ws.Kzx, Mq, nq, Mq, sc[11:12], a_tri=ops.TRI_UPPER, chunk=H),
 "G = A_g A^T (lower, split-K)": lambda: ops.gemm_tch((ws.Agh, ws.Agl), (ws.Ah, ws.Al), ws.G, Mq, Mq, nq, sc[12:13], b_kmajor=True, c_lower=True, chunk=H, nsplit=ws.syrk_split, split_ws=ws.split_ws),
}
res = {}
for tile in (0, 1):
    ops.set_tc_persistent(tile)
    out = {}
    for name, fn in prods.items():
        ms = t(fn)
        tgt = {"A": ws.A, "C": ws.C, "d": ws.Kzx, "G": ws.G}[name[0]]
        out[name] = (ms, tgt[:, : (nq if name[0] != "G" else Mq)].clone())
    ms_step = t(lambda: arm.step(x, V, y), 10)
    res[tile] = (out, ms_step)
    print(f"persistent={tile}: step {ms_step:.3f} ms | " + " | ".join(f"{k.split(' ')[0]} {v[0]:.3f}" for k, v in out.items()), flush=True)
for name in prods:
    a, b = res[0][0][name][1], res[1][0][name][1]
    if name[0] == "G": a, b = a.tril(), b.tril()
    print(name, "finite:", bool(torch.isfinite(a).all()), "bit-identical:", bool(torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))))
ops.set_tc_persistent(1)
