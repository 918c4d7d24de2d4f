"""Whole training iteration (gather + fused step + the two optimisers + schedulers): FusedAdam vs torch.optim.Adam."""
import sys, os, random, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import gp
from dsvgp_b200.optim import FusedAdam
from dsvgp_b200.data import DeviceMinibatchSampler
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
res = {}
for n in (int(a) for a in sys.argv[2:] or ["16384", "512"]):
    wl = dict(bench.WORKLOADS[name], n=n); dtype = torch.float32; dev = torch.device("cuda", 0)
    d, p = wl["d"], wl["p"]
    g = torch.Generator().manual_seed(0)
    X = torch.rand(200000, d, generator=g); Y = torch.randn(200000, d + 1, generator=g)
    sampler = DeviceMinibatchSampler(X, Y, n, dev)
    for kind in ("fused", "torch"):
        model, lik = bench.build_model(wl, dtype, dev)
        mll = gp.VariationalELBO(lik, model, num_data=(d + 1) * 200000)
        vd = model.variational_strategy._variational_distribution
        if kind == "fused":
            vo = FusedAdam([{"params": model.variational_parameters()}], lr=0.01, lower_triangular=[vd.chol_variational_covar])
            ho = FusedAdam([{"params": model.hyperparameters()}, {"params": lik.parameters()}], lr=0.01)
        else:
            vo = torch.optim.Adam([{"params": model.variational_parameters()}], lr=0.01)
            ho = torch.optim.Adam([{"params": model.hyperparameters()}, {"params": lik.parameters()}], lr=0.01)
        vs_, hs_ = (torch.optim.lr_scheduler.MultiStepLR(o, [1000, 2000], gamma=0.1) for o in (vo, ho))
        it = iter(sampler.epoch())
        def iteration():
            idx = next(it)
            x, y, V = sampler.gather(idx, DeviceMinibatchSampler.draw_columns(p, d))
            vo.zero_grad(); ho.zero_grad()
            loss = -mll(lik(model(x, derivative_directions=V)), y)
            loss.backward()
            vo.step(); vs_.step(); ho.step(); hs_.step()
        for _ in range(3): iteration()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): iteration()
        e1.record(); torch.cuda.synchronize()
        res[f"{name}_n{n}_{kind}_ms_per_iteration"] = e0.elapsed_time(e1) / 8
print(json.dumps(res, indent=1))
