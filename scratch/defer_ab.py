"""A/B of engine.DEFER_SIDE_WORK (side-stream work released at the mid-chain diagonal block) on the training step: run(name, n, mids)
times the step for each release point (None = no deferral).  As committed it loops over the lookahead / priority knobs at the
default release point (the last comparison made with it: no gain from either with the 31 us diagonal block)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import gp, ops, engine
dev = torch.device("cuda", 0)
def run(name, n, mids):
    wl = dict(bench.WORKLOADS[name]); dtype = torch.float64 if wl["dtype"] == "f64" else torch.float32
    model, lik = bench.build_model(wl, dtype, dev)
    mll = gp.VariationalELBO(lik, model, num_data=(wl["d"] + 1) * wl["N"])
    x, V, y = (t.to(dev) for t in bench.synth_batch(n, wl["d"], wl["p"], wl["variant"], dtype, "cpu", 1000))
    params = list(model.parameters()) + list(lik.parameters())
    def step():
        for q in params: q.grad = None
        loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward(); return loss
    def t(reps=20):
        for _ in range(4): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): step()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    out = []
    for rnd in range(2):
        for mid in mids:
            engine.DEFER_SIDE_WORK = mid is not None
            if mid is not None: ops.set_chol_mid_link(mid)
            out.append((mid, round(t(), 3)))
    print(name, n, out, flush=True)
for la, pr in ((0, 0), (1, 0), (0, 1), (1, 1)):
    ops.set_chol_lookahead(la); ops.set_chol_priority(pr)
    print("lookahead", la, "priority", pr)
    run("C3", 16384, (20,))
    run("C3", 512, (20,))
    run("C4", 2048, (20,))
ops.set_chol_lookahead(0); ops.set_chol_priority(0)
