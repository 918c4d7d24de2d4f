"""GPU experiment (not a test): which fp32 stage drives the gradient error of the fp32 step?"""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch
from oracle import dsvgp_oracle as O
from dsvgp_b200 import ops, engine
F32, F64 = torch.float32, torch.float64
def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / b.abs().max())
orig_gemm = ops.gemm
state = {"i": 0, "up": set()}
def patched(A, B, C, **kw):
    i = state["i"]; state["i"] += 1
    if C.dtype == F32 and (i in state["up"] or "all" in state["up"]):
        A64, B64, C64 = A.double(), B.double(), C.double()
        D = kw.get("D"); kw2 = dict(kw)
        if D is not None: kw2["D"] = D.double()
        orig_gemm(A64, B64, C64, **kw2)
        C.copy_(C64.float()); return C
    return orig_gemm(A, B, C, **kw)
engine.ops.gemm = patched
for (variant, n, d, M, p, seed) in [("dsvgp", 200, 2, 20, 2, 222), ("dfree", 257, 18, 64, 2, 339)]:
    P, x, Vx, y, nd = O.make_problem(n, d, M, p, F32, seed=seed, variant=variant, N=10 * n)
    P64 = P.clone(F64); up = lambda t: None if t is None else t.double()
    rv, rg = O.elbo_and_grads(P64, up(x), up(Vx), up(y), nd, variant)
    Pg = types.SimpleNamespace(**{k: (v.cuda().contiguous()) for k, v in P.tensors().items()})
    p2 = 0 if variant == "dfree" else p
    print(variant, n, d, M, p)
    for label, ups in [("none", set()), ("all", {"all"})] + [(f"gemm{i}", {i}) for i in range(11)] + [("0+3", {0,3}), ("0,3,5,6", {0,3,5,6})]:
        state["i"] = 0; state["up"] = ups
        elbo, g, mu, var = engine.ENGINE.elbo_step(Pg, x.cuda(), None if Vx is None else Vx.cuda(), y.cuda(), nd, p, p2)
        errs = {k: rel(g[k], rg[k]) for k in ("Z", "Vz", "m", "Ls_raw", "raw_ell", "raw_os", "raw_noise")}
        print("  %-10s elbo %.1e " % (label, abs(float(elbo) - float(rv)) / abs(float(rv))), " ".join("%s %.1e" % kv for kv in errs.items()), "ngemm", state["i"])
