"""cProfile of the host side of end-to-end steps (where the Python time before the first launches goes)."""
import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS["C3"])
arm = bench.Arm(wl, dev, 0, 1)
xh, Vh, yh = arm.batch(wl["n"], 1000)
def e2e_step():
    xb, Vb, yb = xh.to(dev, non_blocking=True), Vh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)
    return float(arm.step(xb, Vb, yb).item())
for _ in range(5): e2e_step()
pr = cProfile.Profile()
pr.enable()
for _ in range(30): e2e_step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
