"""K_zx assembly alone (what RBFKernelDirectionalGrad.forward returns: the fp32 matrix), for an ncu capture."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import gp, ops
from dsvgp_b200.engine import ENGINE
wl = dict(bench.WORKLOADS["C3"]); dtype = torch.float32; dev = torch.device("cuda", 0)
model, lik = bench.build_model(wl, dtype, dev)
n, d, M, p = wl["n"], wl["d"], wl["M"], wl["p"]
mll = gp.VariationalELBO(lik, model, num_data=(d + 1) * wl["N"])
x, V, y = (t.to(dev) for t in bench.synth_batch(n, d, p, "dsvgp", dtype, "cpu", 1000))
loss = -mll(lik(model(x, derivative_directions=V)), y); loss.backward()
ws = ENGINE.workspace(dev, dtype, n, d, M, p, p); f = ENGINE.factor(dev, dtype, d, M, p)
Z = model.variational_strategy.inducing_points.detach(); wx = ops.normalize_dirs(V, dtype)[0]
for _ in range(2): ops.kdir_fwd(Z, f.uzT, p, x, wx, p, f.hyp, ws.Kzx, canon=ws.canon)
torch.cuda.synchronize(); torch.cuda.profiler.start()
ops.kdir_fwd(Z, f.uzT, p, x, wx, p, f.hyp, ws.Kzx, canon=ws.canon)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
