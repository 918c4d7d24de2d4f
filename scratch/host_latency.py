"""Host-side time line of one end-to-end step up to the point where the factorisation is enqueued (perf_counter stamps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from dsvgp_b200 import engine, ops
from dsvgp_b200.engine import Engine, ENGINE
dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS["C3"])
arm = bench.Arm(wl, dev, 0, 1)
xh, Vh, yh = arm.batch(wl["n"], 1000)
stamps = {}
def wrap(obj, name, label, static=False):
    fn = getattr(obj, name)
    def w(*a, **k):
        stamps.setdefault(label + ":in", time.perf_counter())
        r = fn(*a, **k)
        stamps.setdefault(label + ":out", time.perf_counter())
        return r
    setattr(obj, name, staticmethod(w) if static else w)
wrap(ENGINE, "elbo_step", "elbo_step")
wrap(Engine, "_validate", "validate", static=True)
wrap(ENGINE, "_data_dirs", "data_dirs")
wrap(Engine, "_prep", "prep", static=True)
wrap(Engine, "_scales0", "scales0", static=True)
wrap(Engine, "_factorise", "factorise", static=True)
wrap(ops, "cholesky_inverse", "chol_enqueue")
wrap(Engine, "_assemble", "assemble", static=True)
wrap(Engine, "_check", "check", static=True)
def e2e_step():
    stamps.clear()
    t0 = time.perf_counter()
    xb, Vb, yb = xh.to(dev, non_blocking=True), Vh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)
    stamps["copies:out"] = time.perf_counter()
    loss = arm.step(xb, Vb, yb)
    stamps["step_enqueued"] = time.perf_counter()
    v = float(loss.item())
    stamps["loss_read"] = time.perf_counter()
    return t0
for _ in range(5): e2e_step()
acc = {}
for _ in range(10):
    t0 = e2e_step()
    for k, v in stamps.items(): acc.setdefault(k, []).append((v - t0) * 1e6)
for k, v in sorted(acc.items(), key=lambda kv: sum(kv[1])):
    print(f"{k:22s} {sum(v) / len(v):9.1f} us")
