"""Error table of the trained-state cases: GPU paths (3xFP16 / 3xTF32 / mma.sync) and the reference-structured fp32 CPU
oracle, each against the fp64 oracle.  Run on the GPU box:  python scratch/diag_trained.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from oracle import dsvgp_oracle as O
from dsvgp_b200 import engine
import test_step_gpu as T

F32, F64 = torch.float32, torch.float64
CASES = [("dsvgp", 512, 10, 1024, 2, 0.7, "optimal"), ("dsvgp", 512, 10, 1024, 2, 2.0, "optimal"), ("dsvgp", 512, 10, 1024, 2, 0.7, "rough"),
         ("dsvgp", 512, 60, 800, 3, 2.0, "optimal"), ("dfree", 512, 18, 1024, 2, 2.0, "rough"), ("dsvgp", 300, 10, 96, 2, 0.7, "optimal"),
         ("grad", 60, 3, 33, 3, 0.7, "optimal"), ("dsvgp", 512, 10, 1024, 2, 0.7, "near")]
if len(sys.argv) > 1:
    CASES = [CASES[int(a)] for a in sys.argv[1:]]
up = lambda t: None if t is None else t.double()
for (variant, n, d, M, p, ell, kind) in CASES:
    if kind == "near":
        P, x, Vx, y, nd = O.make_problem(n, d, M, p, F32, seed=2, variant=variant, N=100 * n)
    else:
        P, x, Vx, y, nd = O.make_trained_problem(n, d, M, p, F32, seed=2, variant=variant, ell=ell, kind=kind, N=100 * n)
    P64 = P.clone(F64)
    rv, rg = O.elbo_and_grads(P64, up(x), up(Vx), up(y), nd, variant)
    mean, var = O.predict(P64, up(x), up(Vx), variant)
    rows = {}
    t0 = time.time()
    cv, cg = O.elbo_and_grads(P, x, Vx, y, nd, variant, structure="reference")
    cm, cvv = O.predict(P, x, Vx, variant, "reference")
    rows["cpu_ref_fp32"] = (abs(float(cv) - float(rv)) / abs(float(rv)), {k: T.rel(cg[k], rg[k]) for k in rg}, T.rel(cm, mean), T.rel(cvv, var))
    for name, (tc, h, w64) in {"3xFP16": (True, True, False), "3xFP16+W64": (True, True, True), "mma.sync": (False, False, False)}.items():
        engine.USE_TC, engine.USE_FP16, engine.WHITEN_FP64 = tc, h, w64
        engine.ENGINE._ws.clear(), engine.ENGINE._fac.clear()
        model, lik, val, grads, out = T.run_step(variant, P, x, Vx, y, nd, d, F32)
        rows[name] = (abs(float(val) - float(rv)) / abs(float(rv)), {k: T.rel(grads[k], rg[k]) for k in rg}, T.rel(out.mean, mean), T.rel(out.variance, var))
    engine.USE_TC, engine.USE_FP16, engine.WHITEN_FP64 = True, True, False
    print(f"== {variant} n={n} d={d} M={M} p={p} ell={ell} {kind}: elbo {float(rv):.4f} max|m| {float(P.m.abs().max()):.2f}")
    for name, (ev, eg, em, evar) in rows.items():
        print(f"  {name:13s} elbo {ev:.1e} mean {em:.1e} var {evar:.1e} | " + " ".join(f"{k}:{v:.1e}" for k, v in eg.items()))
    sys.stdout.flush()
