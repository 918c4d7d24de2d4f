"""dsvgp_set_chol_graph: the factorisation as a cached CUDA graph.  Bit-identity against plain launches, then the step with and
without it -- device-timed back-to-back steps and a loop that reads the loss back every step (host not running ahead)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import gp, ops
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
F64 = torch.float64
Mq = 3072
g = torch.Generator().manual_seed(1)
R = torch.randn(Mq, Mq + 5, generator=g, dtype=F64)
A = (R @ R.T / (Mq + 5) + 1e-3 * torch.eye(Mq, dtype=F64)).cuda()
Mp, nb0, nlev = ops.chol_plan(Mq)
Aw, L, W = (torch.empty(Mp, Mp, dtype=F64, device=dev) for _ in range(3))
info = torch.ones(1, dtype=torch.int32, device=dev)
def run():
    Aw.copy_(A); L.fill_(float("nan")); W.fill_(float("nan")); info.fill_(1)
    ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    return L.tril().clone(), W.tril().clone()
L0, W0 = run()
ops.set_chol_graph(1)
for i in range(5):
    L1, W1 = run()
    assert torch.equal(L1, L0) and torch.equal(W1, W0), i
def t_alone(reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for on, ns in ((0, 1), (0, 3), (1, 3), (0, 1), (0, 3), (1, 3)):
    ops.set_chol_graph(on); ops.set_chol_inv_streams(ns); t_alone(3)
    L1, W1 = run(); assert torch.equal(L1, L0) and torch.equal(W1, W0)
    print("factorisation alone, graph", on, "inverse streams", ns, f"{t_alone():.3f} ms", flush=True)
ops.set_chol_graph(0)
print("bit-identical with the graph: ok", flush=True)
for name, n in (("C3", 16384), ("C3", 512)):
    wl = dict(bench.WORKLOADS[name])
    model, lik = bench.build_model(wl, torch.float32, dev)
    mll = gp.VariationalELBO(lik, model, num_data=(wl["d"] + 1) * wl["N"])
    x, V, y = (t.to(dev) for t in bench.synth_batch(n, wl["d"], wl["p"], wl["variant"], torch.float32, "cpu", 1000))
    params = list(model.parameters()) + list(lik.parameters())
    def step():
        for q in params: q.grad = None
        loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward(); return loss
    def flat(): return torch.cat([q.grad.reshape(-1).double() for q in params])
    out = []
    ref = None
    for rnd in range(2):
        for on in (1, 3):
            ops.set_chol_inv_streams(on)
            for _ in range(4): step()
            gr = flat().clone()
            if ref is None: ref = gr
            assert torch.equal(gr, ref), "gradients differ"
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): step()
            e1.record(); torch.cuda.synchronize()
            dev_ms = e0.elapsed_time(e1) / 20
            t0 = time.perf_counter()
            for _ in range(20): float(step())
            sync_ms = (time.perf_counter() - t0) / 20 * 1e3
            out.append((on, round(dev_ms, 3), round(sync_ms, 3)))
    print(name, n, "(inverse streams, back-to-back ms, loss-read-every-step ms):", out, flush=True)
ops.set_chol_graph(0)
