"""Timeline of the blocked Cholesky + inverse alone (torch.profiler, CUDA activities): per-kernel durations and the
start-to-start spacing of the diagonal-block kernels, i.e. what the critical path really is.
    python scratch/chol_trace.py [Mq] [variant]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
from dsvgp_b200 import ops, _lib
Mq = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 3
_lib.call_raw("dsvgp_set_chol_variant", variant)
F64 = torch.float64
g = torch.Generator().manual_seed(Mq)
R = torch.randn(Mq, Mq + 5, generator=g, dtype=F64)
A = (R @ R.T / (Mq + 5) + 1e-3 * torch.eye(Mq, dtype=F64)).cuda()
Mp, nb0, nlev = ops.chol_plan(Mq)
Aw = torch.empty(Mp, Mp, dtype=F64, device="cuda"); L = torch.empty_like(Aw); W = torch.empty_like(Aw)
info = torch.ones(1, dtype=torch.int32, device="cuda")
def run():
    Aw.zero_(); Aw[:Mq, :Mq] = A; ops.pad_identity(Aw, Mq)
    ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
rows = [(e.time_range.start - t0, e.time_range.end - t0, f"s{getattr(e, 'device_resource_id', -1)} " + e.name[:40]) for e in ev]
print(f"Mq={Mq} variant={variant}: {len(rows)} kernels, span {rows[-1][1] / 1e3:.3f} ms")
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in rows:
    agg[n][0] += 1; agg[n][1] += e - s
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:42s} x{c:3d} total {t / 1e3:7.3f} ms  avg {t / c:7.1f} us")
pot = [(s, e) for s, e, n in rows if "potrf" in n]
print("potrf start-to-start (us):", " ".join(f"{pot[i + 1][0] - pot[i][0]:.0f}" for i in range(len(pot) - 1)))
print("potrf durations (us):", " ".join(f"{e - s:.0f}" for s, e in pot))
print("first 40 events:")
for s, e, n in rows[:40]:
    print(f"  {s:8.1f} {e - s:7.1f}  {n}")
print("last potrf end", pot[-1][1], "total end", rows[-1][1])
print("events from the 28th diagonal block on (start, duration, stream, name):")
t28 = pot[-1][0] if len(pot) > 27 else 0
for s, e, n in rows:
    if s >= t28:
        print(f"  {s:8.1f} {e - s:7.1f}  {n}")
