"""Time the parts of one training step separately with CUDA events (not under a profiler), and sweep the
benchmarking knobs (assembly rows per CTA / evict-first stores).  Usage:
    python scratch/bench_parts.py [C3|C2|...] [n]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import gp, ops, engine
from dsvgp_b200.engine import ENGINE

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
wl = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2: wl["n"] = int(sys.argv[2])
dtype = torch.float64 if wl["dtype"] == "f64" else torch.float32
dev = torch.device("cuda", 0)
model, lik = bench.build_model(wl, dtype, dev)
n, d, M, p = wl["n"], wl["d"], wl["M"], wl["p"]
p2 = 0 if wl["variant"] == "dfree" else p
mll = gp.VariationalELBO(lik, model, num_data=(d + 1) * wl["N"])
x, V, y = (t.to(dev) for t in bench.synth_batch(n, d, p, wl["variant"], dtype, "cpu", 1000))
params = list(model.parameters()) + list(lik.parameters())

def step():
    for q in params: q.grad = None
    loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward(); return loss

def timed(fn, k=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

res = {}
for _ in range(3): step()
res["step_ms"] = timed(step, 10)
if dtype == torch.float32:                # A/B in one process: one dense (S - I) A product vs the two triangular ones
    for flag in (False, True, False, True):
        engine.DENSE_D = flag
        res[f"step_ms_dense_d_{int(flag)}" + ("" if f"step_ms_dense_d_{int(flag)}" not in res else "_again")] = timed(step, 10)
    engine.DENSE_D = True
ws = ENGINE.workspace(dev, dtype, n, d, M, p, p2)
f = ENGINE.factor(dev, dtype, d, M, p)
vs = model.variational_strategy
P = vs._params() if hasattr(vs, "_params") else None
Mq, nq = ws.Mq, ws.nq

# ---- Cholesky + inverse alone
Kzz0 = torch.empty_like(f.Kzz)
if f.Mp > f.Mq: ops.pad_identity(Kzz0, f.Mq)
ops.kdir_fwd(vs.inducing_points.detach(), f.uz64, f.p, vs.inducing_points.detach(), f.uz64, f.p, f.hyp, Kzz0, diag_add=1e-3)
def chol():
    f.Kzz.copy_(Kzz0)
    ops.cholesky_inverse(f.Kzz, f.L, f.W, f.nb0, f.nlev, f.info)
res["copy_kzz_ms"] = timed(lambda: f.Kzz.copy_(Kzz0))
res["chol_inv_ms"] = timed(chol)
assert int(f.info.item()) == 0
# ---- fp64 tail GEMMs
W, L = f.W, f.L
def tail():
    ops.gemm(W, ws.Xd, ws.dL, ta=True, a_tri=ops.TRI_UPPER, alpha=-1.0, c_tri=1, M=Mq, N=Mq, K=Mq)
    ops.gemm(L, ws.dL, ws.Y, ta=True, a_tri=ops.TRI_UPPER, b_tri=ops.TRI_LOWER, c_tri=1, M=Mq, N=Mq, K=Mq)
    ops.phi_lower(ws.Y, ws.Psi, Mq)
    ops.gemm(ws.Psi, W, ws.Y, a_tri=ops.TRI_LOWER, b_tri=ops.TRI_LOWER, c_tri=1, M=Mq, N=Mq, K=Mq)
    ops.gemm(W, ws.Y, ws.S, ta=True, a_tri=ops.TRI_UPPER, b_tri=ops.TRI_LOWER, M=Mq, N=Mq, K=Mq)
res["tail_ms"] = timed(tail)
if dtype == torch.float64:
    res["whiten_f64_ms"] = timed(lambda: ops.gemm(f.Wt, ws.Kzx, ws.A, a_tri=ops.TRI_LOWER, M=Mq, N=nq, K=Mq))
# ---- assembly knobs
if dtype == torch.float32 and p2:
    Z = vs.inducing_points.detach()
    wx = ops.normalize_dirs(V, dtype)[0]
    for tib in (64, 32, 16):
        for ss in (0, 1):
            ops.set_kdir_fwd_knobs(tib, ss)
            res[f"asm_k_ms_tib{tib}_cs{ss}"] = timed(lambda: ops.kdir_fwd(Z, f.uzT, p, x, wx, p2, f.hyp, ws.Kzx, canon=ws.canon), 20)
            if getattr(ws, "tch", False):
                res[f"asm_half_ms_tib{tib}_cs{ss}"] = timed(lambda: ops.kdir_fwd_half(Z, f.uzT, p, x, wx, p2, f.hyp, ws.Kzx, ws.Kh, ws.Kl, f.scales[1:2], canon=ws.canon), 20)
            else:
                res[f"asm_klo_ms_tib{tib}_cs{ss}"] = timed(lambda: ops.kdir_fwd(Z, f.uzT, p, x, wx, p2, f.hyp, ws.Kzx, canon=ws.canon, out_lo=ws.lo1), 20)
            res[f"asm_gen_ms_tib{tib}_cs{ss}"] = timed(lambda: ops.kdir_fwd(Z, f.uzT, p, x, wx, p2, f.hyp, ws.Kzx), 20)
    ops.set_kdir_fwd_knobs(0, 2)
res["asm_bytes"] = 4 * (Mq * nq + (M + n) * d + (M * p + n * p2) * d)
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"parts_{name}_{n}.json"), "w"), indent=1)
