"""Race detector: the step is deterministic, so repeated identical steps must give bit-identical results (the side
streams of the engine and of the Cholesky would show up here as occasional differences).  Then 150 training iterations
with the fused optimiser: finite losses, no drift in memory."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import gp
from dsvgp_b200.optim import FusedAdam
for name, n in (("C3", 16384), ("C3", 512), ("C5", 4096), ("C2", 2048)):
    wl = dict(bench.WORKLOADS[name], n=n); dtype = torch.float64 if wl["dtype"] == "f64" else torch.float32
    dev = torch.device("cuda", 0)
    model, lik = bench.build_model(wl, dtype, dev)
    mll = gp.VariationalELBO(lik, model, num_data=(wl["d"] + 1) * wl["N"])
    x, V, y = (t.to(dev) for t in bench.synth_batch(n, wl["d"], wl["p"], wl["variant"], dtype, "cpu", 7))
    params = list(model.parameters()) + list(lik.parameters())
    def step():
        for q in params: q.grad = None
        loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward()
        return torch.cat([loss.detach().reshape(1).double()] + [q.grad.reshape(-1).double() for q in params])
    ref = step(); bad = 0
    for i in range(60):
        if not torch.equal(step(), ref): bad += 1
    print(f"{name} n={n}: {bad} of 60 repeated steps differ bitwise from the first", flush=True)
    vd = model.variational_strategy._variational_distribution
    opt = FusedAdam([{"params": params}], lr=0.01, lower_triangular=[vd.chol_variational_covar])
    torch.cuda.reset_peak_memory_stats(); m0 = torch.cuda.memory_allocated()
    losses = []
    for it in range(150):
        opt.zero_grad()
        loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward(); opt.step()
        if it % 50 == 49: losses.append(float(loss.detach()))
    print(f"   150 iterations: losses {losses}, allocated {m0/2**20:.0f} -> {torch.cuda.memory_allocated()/2**20:.0f} MiB, peak {torch.cuda.max_memory_allocated()/2**20:.0f} MiB", flush=True)
    del model, lik, mll, opt; gp.ENGINE._ws.clear(); gp.ENGINE._fac.clear(); torch.cuda.empty_cache()
