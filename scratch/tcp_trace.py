"""Per-item timeline of the persistent 3xFP16 product (dsvgp_set_tc_trace): where the time of a work item goes -- MMA issue
span, wait for the first chunk, chunk-add phase, store phase, and the gap between consecutive items of a CTA pair."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, bench
from dsvgp_b200 import ops, engine, _lib
from dsvgp_b200.engine import ENGINE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
which = sys.argv[2] if len(sys.argv) > 2 else "A"
wl = dict(bench.WORKLOADS["C3"], n=n)
dev = torch.device("cuda", 0)
arm = bench.Arm(wl, dev, 0, 1)
x, V, y = (t.to(dev) for t in arm.batch(n, 1))
for _ in range(3): arm.step(x, V, y)
torch.cuda.synchronize()
ws = ENGINE.workspace(dev, torch.float32, n, wl["d"], wl["M"], wl["p"], wl["p"])
f = ENGINE.factor(dev, torch.float32, wl["d"], wl["M"], wl["p"])
Mq, nq, sc, H = ws.Mq, ws.nq, f.scales, engine.TCH_CHUNK
prods = {
 "A": (lambda: ops.gemm_tch((f.Wh, f.Wl), (ws.Kh, ws.Kl), ws.A, Mq, nq, Mq, sc[8:9], a_tri=ops.TRI_LOWER, chunk=H, Ch=(ws.Ah, ws.Al), c_scale=sc[3:4]), (Mq, nq, Mq, 1, 0, 1)),
 "C": (lambda: ops.gemm_tch((ws.Dh, ws.Dl), (ws.Ah, ws.Al), ws.C, Mq, nq, Mq, sc[13:14], chunk=H), (Mq, nq, Mq, 0, 0, 1)),
 "G": (lambda: ops.gemm_tch((ws.Agh, ws.Agl), (ws.Ah, ws.Al), ws.G, Mq, Mq, nq, sc[12:13], b_kmajor=True, c_lower=True, chunk=H, nsplit=ws.syrk_split, split_ws=ws.split_ws), (Mq, Mq, nq, 0, 1, ws.syrk_split)),
}
fn, (M, N, K, a_tri, c_lower, nz) = prods[which]
P = torch.cuda.get_device_properties(0).multi_processor_count // 2
size = _lib._lib.dsvgp_tc_work_list(M, N, K, a_tri, c_lower, nz, P, 64, None, 0)
buf = np.zeros(size, dtype=np.int32)
_lib._lib.dsvgp_tc_work_list(M, N, K, a_tri, c_lower, nz, P, 64, ctypes.c_void_p(buf.ctypes.data), size)
n_off = (P + 1 + 3) & ~3
offs, items = buf[:P + 1], buf[n_off:].reshape(-1, 4)
nitems = len(items)
fn(); torch.cuda.synchronize()
tr = torch.zeros(nitems * 8, dtype=torch.int64, device=dev)
_lib.call_raw("dsvgp_set_tc_trace", tr, nitems)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
_lib.call_raw("dsvgp_set_tc_trace", None, 0)
ms = e0.elapsed_time(e1)
t = tr.cpu().numpy().reshape(-1, 8).astype(np.float64)
nk = ((items[:, 3].astype(np.int64) >> 16) & 0xFFFF) - (items[:, 3] & 0xFFFF)
ghz = 1.92
us = lambda c: c / (ghz * 1e3)
print(f"product {which}: {ms:.3f} ms, {nitems} items on {P} pairs (clock stamps converted at {ghz} GHz)")
for q in (0, P // 2, P - 1):
    a, b = offs[q], offs[q + 1]
    print(f"pair {q}: {b - a} items, {nk[a:b].sum()} k-blocks; span {us(t[b - 1, 6] - t[a, 0]):.1f} us")
    for i in range(a, min(b, a + 6)):
        gap = us(t[i, 1] - t[i - 1, 2]) if i > a else float("nan")
        print(f"   item {i - a}: nk {nk[i]:3d} | mma issue span {us(t[i, 2] - t[i, 1]):6.2f} ({us(t[i,2]-t[i,1])/max(nk[i],1):.3f}/kb) wait-operands {us(t[i, 1] - t[i, 0]):5.2f} gap-from-prev-item {gap:5.2f}"
              f" | epi: wait first chunk {us(t[i, 4] - t[i, 3]):5.2f}, chunk phase {us(t[i, 5] - t[i, 4]):6.2f}, store phase {us(t[i, 6] - t[i, 5]):5.2f}")
sel = np.arange(nitems)
first = np.zeros(nitems, bool); first[offs[:-1][np.diff(offs) > 0]] = True
store = us(t[:, 6] - t[:, 5]); chunk = us(t[:, 5] - t[:, 4]); issue = us(t[:, 2] - t[:, 1])
gapm = us(t[1:, 1] - t[:-1, 2])[~first[1:]]
print(f"mean store phase {store.mean():.2f} us (min {store.min():.2f}, max {store.max():.2f}); mean chunk phase per k-block {(chunk / np.maximum(nk, 1)).mean():.3f} us;"
      f" mean MMA issue per k-block {(issue / np.maximum(nk, 1)).mean():.3f} us; mean MMA gap between items {gapm.mean():.2f} us")
tot_span = np.array([us(t[offs[q + 1] - 1, 6] - t[offs[q], 0]) for q in range(P) if offs[q + 1] > offs[q]])
print(f"pair spans: min {tot_span.min():.1f} mean {tot_span.mean():.1f} max {tot_span.max():.1f} us")
