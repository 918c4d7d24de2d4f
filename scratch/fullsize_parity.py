"""GPU-box experiment: fp32 engine vs fp64 CPU oracle at FULL inducing sizes (slow oracle, not part of CI)."""
import sys, os, types, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch
from oracle import dsvgp_oracle as O
from dsvgp_b200 import engine
F32, F64 = torch.float32, torch.float64
print("cpu threads", torch.get_num_threads(), os.cpu_count())
def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / b.abs().max())
cases = [("dsvgp", 1024, 10, 1024, 2), ("dsvgp", 512, 60, 800, 3), ("dfree", 1024, 18, 1024, 2), ("dsvgp", 1000, 3, 512, 1)]
if len(sys.argv) > 1: cases = cases[: int(sys.argv[1])]
for (variant, n, d, M, p) in cases:
    P, x, Vx, y, nd = O.make_problem(n, d, M, p, F32, seed=1, variant=variant, N=100 * n)
    P64 = P.clone(F64); up = lambda t: None if t is None else t.double()
    t0 = time.time(); rv, rg = O.elbo_and_grads(P64, up(x), up(Vx), up(y), nd, variant); t1 = time.time()
    Pg = types.SimpleNamespace(**{k: v.cuda().contiguous() for k, v in P.tensors().items()})
    p2 = 0 if variant == "dfree" else p
    elbo, g, mu, var = engine.ENGINE.elbo_step(Pg, x.cuda(), None if Vx is None else Vx.cuda(), y.cuda(), nd, p, p2)
    rm, rvv = O.predict(P64, up(x), up(Vx), variant)
    errs = {k: rel(g[k], rg[k]) for k in ("Z", "Vz", "m", "Ls_raw", "c", "raw_ell", "raw_os", "raw_noise")}
    print(variant, n, d, M, p, "oracle %.1fs" % (t1 - t0), "elbo %.1e" % (abs(float(elbo) - float(rv)) / abs(float(rv))),
          "mean %.1e var %.1e" % (rel(mu, rm), rel(var, rvv)), " ".join("%s %.1e" % kv for kv in errs.items()), flush=True)
