"""One profiled training step of a bench workload (run under ncu --profile-from-start off)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import gp
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
wl = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2: wl["n"] = int(sys.argv[2])
dtype = torch.float64 if wl["dtype"] == "f64" else torch.float32
dev = torch.device("cuda", 0)
model, lik = bench.build_model(wl, dtype, dev)
mll = gp.VariationalELBO(lik, model, num_data=(wl["d"] + 1) * wl["N"])
x, V, y = (t.to(dev) for t in bench.synth_batch(wl["n"], wl["d"], wl["p"], wl["variant"], dtype, "cpu", 1000))
def step():
    for q in list(model.parameters()) + list(lik.parameters()): q.grad = None
    loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward(); return loss
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
