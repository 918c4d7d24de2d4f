"""SURVEY.md section 8f ranks 2 and 3 on the GPU: the fused Adam step against torch.optim.Adam, and the fused
minibatch gather + select_cols_of_y against the reference's op sequence (directional_vi.py:68-90, :239-241)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu
F32, F64 = torch.float32, torch.float64


@pytest.mark.parametrize("dtype,tol", [(F32, 2e-6), (F64, 1e-13)])
@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(dtype, tol, weight_decay):
    from dsvgp_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(37, 5), (300, 300), (1,), (), (1, 1), (4099,), (64, 10)]
    base = [torch.randn(s, generator=g, dtype=F64) for s in shapes]
    mk = lambda: [torch.nn.Parameter(b.to(dtype).cuda()) for b in base]
    pa, pb = mk(), mk()
    with torch.no_grad():                                  # parameter 1 plays chol_variational_covar
        pa[1].copy_(pa[1].tril()), pb[1].copy_(pb[1].tril())
    ref = torch.optim.Adam([{"params": pa[:2]}, {"params": pa[2:], "lr": 0.03}], lr=0.01, weight_decay=weight_decay,
                           foreach=False)
    opt = FusedAdam([{"params": pb[:2]}, {"params": pb[2:], "lr": 0.03}], lr=0.01, weight_decay=weight_decay,
                    lower_triangular=[pb[1]] if weight_decay == 0.0 else [])
    sr = torch.optim.lr_scheduler.MultiStepLR(ref, [3, 6], gamma=0.1)
    so = torch.optim.lr_scheduler.MultiStepLR(opt, [3, 6], gamma=0.1)
    for it in range(9):
        for k, (a, b) in enumerate(zip(pa, pb)):
            gr = torch.randn(a.shape, generator=g, dtype=F64).to(dtype).cuda() * (10.0 ** (k - 3))
            if k == 1:
                gr = gr.tril()
            if k == 5 and it < 2:                          # a tensor that joins late gets its own step count
                a.grad = b.grad = None
                continue
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step(), opt.step()
        sr.step(), so.step()
    for a, b in zip(pa, pb):
        err = float((a.detach().double() - b.detach().double()).abs().max() / a.detach().double().abs().max().clamp_min(1e-30))
        assert err < tol, (tuple(a.shape), err)
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and len(sd["param_groups"]) == 2
    assert float(sd["state"][5]["step"]) == 7.0 and float(sd["state"][0]["step"]) == 9.0


def test_fused_adam_rejects_cpu_parameters():
    from dsvgp_b200 import _lib
    from dsvgp_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(_lib.DsvgpError):
        FusedAdam([p]).step()


@pytest.mark.parametrize("dtype", [F32, F64])
@pytest.mark.parametrize("N,d,p,bs", [(1000, 10, 2, 128), (37, 3, 1, 37), (50, 2, 2, 7), (64, 18, 0, 16)])
def test_gather_batch_matches_reference_ops(dtype, N, d, p, bs):
    import directional_vi
    from dsvgp_b200.data import DeviceMinibatchSampler
    g = torch.Generator().manual_seed(N + d)
    x = torch.rand(N, d, generator=g, dtype=dtype)
    y = torch.randn(N, d + 1, generator=g, dtype=dtype)
    sampler = DeviceMinibatchSampler(x, y, bs)
    assert len(sampler) == (N + bs - 1) // bs
    seen = []
    for idx in sampler.epoch():
        random.seed(len(seen))
        cols = DeviceMinibatchSampler.draw_columns(p, d)
        xb, yb, V = sampler.gather(idx, cols)
        # the reference's sequence on the same rows with the same RNG state
        random.seed(len(seen))
        y_sel, dirs = directional_vi.select_cols_of_y(y[idx.cpu()].cuda(), p, d)
        assert torch.equal(xb.cpu(), x[idx.cpu()])
        assert torch.equal(yb, y_sel.reshape(-1))
        if p:
            assert torch.equal(V, dirs.to(dtype).repeat(len(idx), 1))
        else:
            assert V is None
        seen.append(idx.cpu())
    assert torch.equal(torch.cat(seen).sort().values, torch.arange(N))       # one epoch = every row exactly once


def test_train_gp_runs_with_fused_input_and_optimizer():
    """directional_vi.train_gp end to end on the shipped test problem shape (tests/test_dsvgp.py:21-29): the fused
    gather feeds the fused step, the fused Adam updates every parameter, the loss goes down."""
    import math
    import directional_vi
    torch.manual_seed(0), random.seed(0)
    n, d = 600, 2
    x = torch.rand(n, d)
    f = torch.sin(2 * math.pi * (x ** 2).sum(1, keepdim=True))
    df = 4 * math.pi * x * torch.cos(2 * math.pi * (x ** 2).sum(1, keepdim=True))
    ds = torch.utils.data.TensorDataset(x, torch.cat([f, df], 1))
    losses = []
    import builtins
    real_print = builtins.print
    def spy(*a, **k):
        s = " ".join(str(t) for t in a)
        if "loss:" in s and "total_step" in s:
            losses.append(float(s.split("loss:")[1].split(",")[0]))
    builtins.print = spy
    try:
        model, lik = directional_vi.train_gp(ds, num_inducing=20, num_directions=2, minibatch_size=200, minibatch_dim=2,
                                             num_epochs=60, learning_rate_hypers=0.01, lr_sched="step_lr", verbose=True)
    finally:
        builtins.print = real_print
    assert len(losses) >= 3 and losses[-1] < losses[0] - 0.1, losses
    means, variances = directional_vi.eval_gp(ds, model, lik, num_directions=2, minibatch_size=300, minibatch_dim=2)
    assert means.shape == (n * 3,) and bool((variances > 0).all()) and bool(torch.isfinite(means).all())
