"""GPU idle time between consecutive END-TO-END steps (host copies in, loss read back every step): how much of the e2e - device
difference is the host getting the next step's first kernels out."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS["C3"])
arm = bench.Arm(wl, dev, 0, 1)
xh, Vh, yh = arm.batch(wl["n"], 1000)
def e2e_step():
    xb, Vb, yb = xh.to(dev, non_blocking=True), Vh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)
    return float(arm.step(xb, Vb, yb).item())
for _ in range(4): e2e_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): e2e_step()
print(f"e2e wall per step {(time.perf_counter() - t0) * 100:.3f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4): e2e_step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
firsts = [i for i, e in enumerate(evs) if "Memcpy HtoD" in e.name]
# steps start with 3 HtoD copies
starts = [i for k, i in enumerate(firsts) if k % 3 == 0]
for a, b in zip(starts[1:-1], starts[2:]):
    prev_end = max(e.time_range.end for e in evs[:a])
    first_k = next(e for e in evs[a:] if "Memcpy" not in e.name)
    potrf0 = next(e for e in evs[a:] if "potrf" in e.name)
    print(f"step: GPU idle before the first copy {evs[a].time_range.start - prev_end:7.1f} us; first copy -> first kernel {first_k.time_range.start - evs[a].time_range.start:7.1f} us;"
          f" first copy -> first diagonal block {potrf0.time_range.start - evs[a].time_range.start:7.1f} us; span {(evs[b].time_range.start - evs[a].time_range.start) / 1000:.3f} ms")
