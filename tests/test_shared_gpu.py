"""GPU parity of the shared-direction strategy (second half) through the reference-facing
API: shared_directional_vi.GPModel + likelihood + VariationalELBO against golden vectors of the UNMODIFIED reference
files (SharedDirectionalGradVariationalStrategy.py:95-108 one shared direction set, :209-212 zeroed middle term) and
against the fp64 oracle."""
import os

import pytest
import torch

from oracle import dsvgp_oracle as O
from test_step_gpu import F32, F64, rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def build_shared(P, d, dtype):
    import shared_directional_vi
    from dsvgp_b200 import gp
    cpu = lambda t: t.detach().to(dtype).cpu()
    model = shared_directional_vi.GPModel(cpu(P.Z), cpu(P.Vz), d).to("cuda", dtype)
    lik = gp.GaussianLikelihood().to("cuda", dtype)
    vs, vd = model.variational_strategy, model.variational_strategy._variational_distribution
    with torch.no_grad():
        vd.variational_mean.copy_(P.m)
        vd.chol_variational_covar.copy_(P.Ls_raw)
        vs.variational_params_initialized.fill_(1)
        model.mean_module.constant.copy_(P.c)
        model.covar_module.raw_outputscale.copy_(P.raw_os)
        model.covar_module.base_kernel.raw_lengthscale.copy_(P.raw_ell)
        lik.noise_covar.raw_noise.copy_(P.raw_noise)
    return model, lik


def grads_shared(model, lik):
    vs, vd = model.variational_strategy, model.variational_strategy._variational_distribution
    return dict(Z=vs.inducing_points.grad, Vz=vs.inducing_directions.grad, m=vd.variational_mean.grad,
                Ls_raw=vd.chol_variational_covar.grad, c=model.mean_module.constant.grad,
                raw_os=model.covar_module.raw_outputscale.grad, raw_ell=model.covar_module.base_kernel.raw_lengthscale.grad,
                raw_noise=lik.noise_covar.raw_noise.grad)


def step(P, x, Vx, y, num_data, d, dtype):
    from dsvgp_b200 import gp
    model, lik = build_shared(P, d, dtype)
    model.train(), lik.train()
    mll = gp.VariationalELBO(lik, model, num_data=num_data)
    out = lik(model(x.to(dtype).cuda(), derivative_directions=Vx.to(dtype)))
    loss = -mll(out, y.to(dtype).cuda())
    loss.backward()
    return model, lik, -loss.detach(), {k: -g for k, g in grads_shared(model, lik).items()}, out


@pytest.mark.parametrize("name", ["shared_d3_p2_f64", "shared_d3_p2_f32", "shared_d5_p1_f64", "shared_d6_p3_f32"])
def test_shared_step_matches_unmodified_reference_golden(name):
    c = torch.load(os.path.join(GOLD, "shared_cases.pt"))[name]
    dtype, f64 = c["x"].dtype, c["x"].dtype == F64
    P = O.Params(**c["params"])
    model, lik, val, grads, out = step(P, c["x"], c["Vx"], c["y"], c["num_data"], c["d"], dtype)
    t = 1e-10 if f64 else 1e-4
    assert abs(float(val) - float(c["elbo"])) <= t * abs(float(c["elbo"]))
    for k, g in c["grads"].items():
        assert rel(grads[k], g) < (1e-9 if f64 else 2e-4), (k, rel(grads[k], g))     # fp32 fixture = the reference's own fp32
    assert rel(out.mean, c["train_mean"]) < t and rel(out.variance, c["train_variance"]) < t
    if not f64:
        up = lambda q: q.double()
        rv, rg = O.elbo_and_grads(P.clone(F64), up(c["x"]), up(c["Vx"]), up(c["y"]), c["num_data"], "shared")
        assert abs(float(val) - float(rv)) <= 1e-4 * abs(float(rv))
        for k, g in rg.items():
            assert rel(grads[k], g) < 1e-4, k
    model.eval(), lik.eval()
    with torch.no_grad():
        for _ in range(2):
            preds = lik(model(c["x"].cuda(), derivative_directions=c["Vx"]))
            assert rel(preds.mean, c["pred_mean"]) < t and rel(preds.variance, c["pred_variance"]) < t
        assert rel(preds.covariance_matrix, c["pred_covariance"]) < t


@pytest.mark.parametrize("n,d,M,p,dtype", [(300, 10, 96, 2, F32), (257, 10, 96, 2, F64), (512, 10, 1024, 2, F32)])
def test_shared_step_matches_oracle(n, d, M, p, dtype):
    """... including BASELINE's full M = 1024 (M' = 3072, tcgen05 path)."""
    P, x, Vx, y, num_data = O.make_shared_problem(n, d, M, p, dtype, seed=n + M, N=50 * n)
    up = lambda q: q.double()
    rv, rg = O.elbo_and_grads(P.clone(F64), up(x), up(Vx), up(y), num_data, "shared")
    model, lik, val, grads, out = step(P, x, Vx, y, num_data, d, dtype)
    f64 = dtype == F64
    assert abs(float(val) - float(rv)) <= (1e-10 if f64 else 1e-4) * abs(float(rv))
    for k, g in rg.items():
        assert rel(grads[k], g) < (1e-9 if f64 else 1e-4), (k, rel(grads[k], g))
    mean, var = O.predict(P.clone(F64), up(x), up(Vx), "shared")
    assert rel(out.mean, mean) < (1e-10 if f64 else 1e-4) and rel(out.variance, var) < (1e-10 if f64 else 1e-4)


def test_shared_train_gp_runs_and_learns():
    import math
    import random
    import shared_directional_vi
    torch.manual_seed(2), random.seed(2)
    x = torch.rand(400, 2)
    ds = torch.utils.data.TensorDataset(x, O.testfun(x.double()).float())
    model, lik = shared_directional_vi.train_gp(ds, num_inducing=16, num_directions=2, minibatch_size=200, minibatch_dim=2,
                                                num_epochs=30, inducing_data_initialization=False, verbose=False)
    vs = model.variational_strategy
    assert vs.inducing_directions.shape == (2, 2) and vs._variational_distribution.variational_mean.shape == (18,)
    means, variances = shared_directional_vi.eval_gp(ds, model, lik, num_directions=2, minibatch_size=400, minibatch_dim=2)
    assert means.shape == (1200,) and bool(torch.isfinite(means).all()) and bool((variances > 0).all())
    assert all(math.isfinite(float(q.abs().sum())) for q in model.parameters())


def test_strategy_properties_and_foreign_arguments():
    """prior_distribution / variational_distribution (DGVS.py:77-87, :222) and the refusal of foreign inducing arguments."""
    import directional_vi
    model = directional_vi.GPModel(torch.rand(6, 3), torch.eye(3)[:2].repeat(6, 1), 3).cuda()
    vs = model.variational_strategy
    prior = vs.prior_distribution
    assert prior.mean.shape == (18,) and float(prior.mean.abs().max()) == 0.0
    assert torch.equal(prior.covariance_matrix, torch.eye(18, device="cuda"))
    q = vs.variational_distribution
    Ls = vs._variational_distribution.chol_variational_covar.tril()
    assert torch.allclose(q.covariance_matrix, Ls @ Ls.T) and q.mean.shape == (18,)
    x = torch.rand(5, 3, device="cuda")
    V = torch.eye(3)[:2].repeat(5, 1)
    with pytest.raises(NotImplementedError):
        vs.forward(x, torch.rand(6, 3, device="cuda"), None, derivative_directions=V)
    with pytest.raises(NotImplementedError):
        vs.forward(x, vs.inducing_points, torch.zeros(18, device="cuda"), derivative_directions=V)
    out = vs.forward(x, vs.inducing_points, vs._variational_distribution.variational_mean, derivative_directions=V)
    assert out.mean.shape == (15,)
