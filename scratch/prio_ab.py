"""A/B of dsvgp_set_chol_lookahead / dsvgp_set_chol_priority on the step of several workloads."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from dsvgp_b200 import ops
dev = torch.device("cuda", 0)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, n in (("C3", 16384), ("C3", 512), ("C5", 16384), ("C2", 4096)):
    wl = dict(bench.WORKLOADS[name], n=n)
    arm = bench.Arm(wl, dev, 0, 1)
    x, V, y = (t_.to(dev) for t_ in arm.batch(n, 1))
    for rnd in range(2):
        for ru, la, pr in ((0, 0, 0), (1, 0, 0), (1, 0, 1), (1, 1, 1)):
            ops.set_rank_update(ru); ops.set_chol_lookahead(la); ops.set_chol_priority(pr)
            print(f"{name} n={n} rank_update={ru} lookahead={la} priority={pr}: step {t(lambda: arm.step(x, V, y)):.3f} ms", flush=True)
ops.set_rank_update(0); ops.set_chol_lookahead(0); ops.set_chol_priority(0)
