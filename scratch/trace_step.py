"""Kernel durations of real (back-to-back) training steps via torch.profiler (CUPTI), no replay/serialisation."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench, collections
from torch.profiler import profile, ProfilerActivity
from dsvgp_b200 import gp, ops
if os.environ.get("CHOL_PRIORITY"): ops.set_chol_priority(int(os.environ["CHOL_PRIORITY"]))
if os.environ.get("CHOL_LOOKAHEAD"): ops.set_chol_lookahead(int(os.environ["CHOL_LOOKAHEAD"]))
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
wl = dict(bench.WORKLOADS[name]); dtype = torch.float64 if wl["dtype"] == "f64" else torch.float32
dev = torch.device("cuda", 0)
model, lik = bench.build_model(wl, dtype, dev)
mll = gp.VariationalELBO(lik, model, num_data=(wl["d"] + 1) * wl["N"])
x, V, y = (t.to(dev) for t in bench.synth_batch(wl["n"], wl["d"], wl["p"], wl["variant"], dtype, "cpu", 1000))
params = list(model.parameters()) + list(lik.parameters())
def step():
    for q in params: q.grad = None
    loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward(); return loss
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict()
for e in evs:
    k = e.name.split("(")[0].split("<")[0][-48:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"wall span per step {(t1 - t0) / 5 / 1000:.3f} ms; sum of kernel times per step {tot / 5 / 1000:.3f} ms")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:50s} {c // 5:5d}/step {t / 5 / 1000:9.3f} ms/step")
# individual tch gemm launches of the last step
g = [e for e in evs if "gemm_tch" in e.name][-4:]
print("tch launches (us):", [round(e.device_time if hasattr(e, "device_time") else e.cuda_time) for e in g])
# timeline of the last step: Cholesky phase and gaps on the critical path
last = [e for e in evs if e.time_range.start >= sorted(e2.time_range.start for e2 in evs if "hyp_from_raw" in e2.name)[-1]]
last.sort(key=lambda e: e.time_range.start)
po = [e for e in last if "potrf" in e.name]
print(f"potrf: n={len(po)} first start -> last end {(po[-1].time_range.end - po[0].time_range.start)/1000:.3f} ms, sum of potrf {sum(e.time_range.end - e.time_range.start for e in po)/1000:.3f} ms")
gaps = [(po[i+1].time_range.start - po[i].time_range.end) for i in range(len(po)-1)]
print("gaps between consecutive potrf (us):", [round(g) for g in gaps])
t_first = last[0].time_range.start
def when(sub, which=0):
    xs = [e for e in last if sub in e.name]
    return None if not xs else (xs[which].time_range.start - t_first, xs[which].time_range.end - t_first)
for nm in ("hyp_from_raw", "kdir_fwd_v4", "kdir_fwd_blocked", "potrf", "split_half", "gemm_tch", "col_dots", "dA_half", "kdir_bwd_v4", "splitk_reduce", "gemm_tc2", "kdir_bwd_blocked", "var_grads"):
    w = when(nm); wl_ = when(nm, -1)
    if w: print(f"{nm:20s} first {w[0]/1000:8.3f} .. last end {wl_[1]/1000:8.3f} ms")
print("step span", (last[-1].time_range.end - t_first)/1000)
# idle gaps (> 8 us with NO kernel running) in the last step and what follows them
iv = sorted((e.time_range.start, e.time_range.end, e.name.split("(")[0].split("<")[0][-40:]) for e in last)
cur_end = iv[0][1]; idle = 0.0
print("idle gaps > 8us:")
for s0, e0, nm in iv[1:]:
    if s0 - cur_end > 8:
        print(f"   at {(s0 - t_first)/1000:7.3f} ms: {s0 - cur_end:6.1f} us before {nm}")
    if s0 > cur_end: idle += s0 - cur_end
    cur_end = max(cur_end, e0)
print(f"total idle inside the step: {idle/1000:.3f} ms")
if len(sys.argv) > 2:
    print("first kernels of the last step (start us, duration us, stream, name):")
    for e in last[: int(sys.argv[2])]:
        print(f"  {(e.time_range.start - t_first):8.1f} {(e.time_range.end - e.time_range.start):7.1f}  s{getattr(e, 'device_index', 0)}:{getattr(e, 'device_resource_id', getattr(e, 'stream', -1))}  {e.name.split('(')[0][-60:]}")
if os.environ.get("TRACE_DUMP"):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", os.environ["TRACE_DUMP"]), "w") as fh:
        prev_end = None
        for e in last:
            s0, e0 = e.time_range.start - t_first, e.time_range.end - t_first
            fh.write(f"{s0:9.1f} {e0 - s0:8.1f} gap {0.0 if prev_end is None else s0 - prev_end:7.1f} s{getattr(e, 'device_resource_id', -1)} {e.name.split('(')[0][-70:]}\n")
            prev_end = e0 if prev_end is None else max(prev_end, e0)
