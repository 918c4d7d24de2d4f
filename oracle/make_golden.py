"""Generate tests/golden/*.pt by running the UNMODIFIED reference files on CPU -- TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs /root/reference, which the GPU box does not have):

    python oracle/make_golden.py

The reference's hot-path files (RBFKernelDirectionalGrad.py, DirectionalGradVariationalStrategy.py,
DFreeDirectionalGradVariationalStrategy.py, GradVariationalStrategy.py, and GPModel from directional_vi.py /
dfree_directional_vi.py / grad_svgp.py) are imported from where they lie under /root/reference/directionalvi
with `gpytorch` resolved to oracle/gpytorch_shim (gpytorch==1.4.0 itself is not installable here).  Inputs come
from oracle.dsvgp_oracle.make_problem, so every fixture carries the exact inputs and the reference's outputs:
kernel matrices, the ELBO value, its gradient w.r.t. every parameter, train-mode output mean / variance and
eval-mode predictions.  No reference source is copied; only numbers are stored.
"""
import contextlib
import io
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/directionalvi"
sys.path.insert(0, os.path.join(HERE, "gpytorch_shim"))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "utils"))
for name in ("matplotlib", "matplotlib.pyplot", "wandb"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

import gpytorch  # noqa: E402  (the shim)
from oracle import dsvgp_oracle as O  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from RBFKernelDirectionalGrad import RBFKernelDirectionalGrad  # noqa: E402  (reference file)
    import directional_vi as ref_dsvgp  # noqa: E402  (reference file)
    import dfree_directional_vi as ref_dfree  # noqa: E402  (reference file)
    import grad_svgp as ref_grad  # noqa: E402  (reference file)
    import shared_directional_vi as ref_shared  # noqa: E402  (reference file; SharedDirectionalGradVariationalStrategy.py)


def kernel_case(n1, n2, d, p, dtype, seed, same=False):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(n1, d, generator=g, dtype=dtype)
    v1 = torch.randn(n1 * p, d, generator=g, dtype=dtype)
    if same:
        x2, v2 = x1, v1
    else:
        x2 = torch.rand(n2, d, generator=g, dtype=dtype)
        v2 = torch.randn(n2 * p, d, generator=g, dtype=dtype)
    raw_ell = torch.tensor([[0.3]], dtype=dtype)
    k = RBFKernelDirectionalGrad().to(dtype)
    k.raw_lengthscale.data.copy_(raw_ell)
    with torch.no_grad():
        K = k(x1, x2, v1=v1, v2=v2).evaluate()
        out = dict(x1=x1, x2=x2, v1=v1, v2=v2, raw_ell=raw_ell, K=K)
        if same:
            out["Kdiag"] = k(x1, x1, diag=True, v1=v1, v2=v1)
    return out


def _load(model, likelihood, P, variant):
    vs = model.variational_strategy
    with torch.no_grad():
        vs.inducing_points.data = P.Z.clone()
        if variant != "grad":
            vs.inducing_directions.data = P.Vz.clone()
        vs._variational_distribution.variational_mean.data = P.m.clone()
        vs._variational_distribution.chol_variational_covar.data = P.Ls_raw.clone()
        vs.variational_params_initialized.fill_(1)
        model.mean_module.constant.data = P.c.clone()
        model.covar_module.raw_outputscale.data = P.raw_os.clone()
        model.covar_module.base_kernel.raw_lengthscale.data = P.raw_ell.clone()
        likelihood.noise_covar.raw_noise.data = P.raw_noise.clone()


def _grads(model, likelihood, variant):
    vs = model.variational_strategy
    g = dict(Z=vs.inducing_points.grad, m=vs._variational_distribution.variational_mean.grad,
             Ls_raw=vs._variational_distribution.chol_variational_covar.grad, c=model.mean_module.constant.grad,
             raw_os=model.covar_module.raw_outputscale.grad,
             raw_ell=model.covar_module.base_kernel.raw_lengthscale.grad, raw_noise=likelihood.noise_covar.raw_noise.grad)
    if variant != "grad":
        g["Vz"] = vs.inducing_directions.grad
    return {k: v.detach().clone() for k, v in g.items()}


def strategy_case(variant, n, d, M, p, dtype, seed, perturb_dirs=True):
    """One training step (loss = -mll(likelihood(model(x)), y); backward) and one eval-mode prediction,
    exactly as train_gp / eval_gp drive them (directional_vi.py:243-249, :296-298)."""
    P, x, Vx, y, num_data = O.make_problem(n, d, M, p, dtype, seed, variant, perturb_dirs)
    torch.set_default_dtype(dtype)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            if variant == "dsvgp":
                model = ref_dsvgp.GPModel(P.Z, P.Vz, d)
            elif variant == "dfree":
                model = ref_dfree.GPModel(P.Z, P.Vz, d)
            else:
                model = ref_grad.GPModel(P.Z)
        likelihood = gpytorch.likelihoods.GaussianLikelihood()
        model, likelihood = model.to(dtype), likelihood.to(dtype)
        _load(model, likelihood, P, variant)
        model.train(), likelihood.train()
        kwargs = {} if variant == "grad" else {"derivative_directions": Vx}
        mll = gpytorch.mlls.VariationalELBO(likelihood, model, num_data=num_data)
        output = likelihood(model(x, **kwargs))
        loss = -mll(output, y)
        loss.backward()
        grads = {k: -v for k, v in _grads(model, likelihood, variant).items()}   # d ELBO / d param
        res = dict(variant=variant, n=n, d=d, M=M, p=p, seed=seed, perturb_dirs=perturb_dirs, num_data=num_data,
                   x=x, Vx=Vx, y=y, params=P.tensors(), elbo=(-loss).detach().clone(), grads=grads,
                   train_mean=output.mean.detach().clone(), train_variance=output.variance.detach().clone())
        # the same step with mll_type="PLL" (directional_vi.py:218-219): PredictiveLogLikelihood on likelihood(model(x))
        for q in list(model.parameters()) + list(likelihood.parameters()):
            q.grad = None
        pll = gpytorch.mlls.PredictiveLogLikelihood(likelihood, model, num_data=num_data)
        val = pll(likelihood(model(x, **kwargs)), y)
        (-val).backward()
        res["pll"] = val.detach().clone()
        res["pll_grads"] = {k: -v for k, v in _grads(model, likelihood, variant).items()}
        model.eval(), likelihood.eval()
        with torch.no_grad():
            preds = likelihood(model(x, **kwargs))
            res["pred_mean"], res["pred_variance"] = preds.mean.clone(), preds.variance.clone()
            res["pred_covariance"] = preds.covariance_matrix.clone()      # what preds.sample factorises (test_turbo.py:138)
        return res
    finally:
        torch.set_default_dtype(torch.float32)


def shared_case(n, d, M, p, dtype, seed):
    """shared_directional_vi.GPModel (one direction set for all inducing points, M + p variational values, middle term
    zeroed -- SharedDirectionalGradVariationalStrategy.py:95-108, :209-212): one training step and one eval prediction."""
    P, x, Vx, y, num_data = O.make_shared_problem(n, d, M, p, dtype, seed)
    torch.set_default_dtype(dtype)        # the reference builds `iv = torch.zeros(...)` in the default dtype (:100)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref_shared.GPModel(P.Z, P.Vz, d)
        likelihood = gpytorch.likelihoods.GaussianLikelihood()
        model, likelihood = model.to(dtype), likelihood.to(dtype)
        _load(model, likelihood, P, "shared")
        model.train(), likelihood.train()
        kwargs = {"derivative_directions": Vx}
        mll = gpytorch.mlls.VariationalELBO(likelihood, model, num_data=num_data)
        output = likelihood(model(x, **kwargs))
        loss = -mll(output, y)
        loss.backward()
        grads = {k: -v for k, v in _grads(model, likelihood, "shared").items()}
        res = dict(variant="shared", n=n, d=d, M=M, p=p, seed=seed, num_data=num_data, x=x, Vx=Vx, y=y, params=P.tensors(),
                   elbo=(-loss).detach().clone(), grads=grads, train_mean=output.mean.detach().clone(),
                   train_variance=output.variance.detach().clone())
        model.eval(), likelihood.eval()
        with torch.no_grad():
            preds = likelihood(model(x, **kwargs))
            res["pred_mean"], res["pred_variance"] = preds.mean.clone(), preds.variance.clone()
            res["pred_covariance"] = preds.covariance_matrix.clone()
        return res
    finally:
        torch.set_default_dtype(torch.float32)


def ngd_case(n, d, M, p, dtype, seed):
    """The reference model with variational_distribution="NGD" (directional_vi.py:38-40): one training step through
    NaturalVariationalDistribution; the gradients of natural_vec / natural_mat are the natural gradients that
    gpytorch.optim.NGD applies (:187, :251).  Also the parameters after one NGD step."""
    P, x, Vx, y, num_data = O.make_problem(n, d, M, p, dtype, seed, "dsvgp", True)
    torch.set_default_dtype(dtype)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref_dsvgp.GPModel(P.Z, P.Vz, d, variational_distribution="NGD")
        likelihood = gpytorch.likelihoods.GaussianLikelihood()
        model, likelihood = model.to(dtype), likelihood.to(dtype)
        vs = model.variational_strategy
        vd = vs._variational_distribution
        g = torch.Generator().manual_seed(seed)
        Mq = M * (p + 1)
        R = torch.randn(Mq, Mq, generator=g, dtype=torch.float64) / Mq ** 0.5
        prec = (torch.eye(Mq, dtype=torch.float64) + 0.3 * R @ R.T).to(dtype)          # S^-1, SPD
        nat_vec = (0.1 * torch.randn(Mq, generator=g, dtype=torch.float64)).to(dtype)
        with torch.no_grad():
            vs.inducing_points.data = P.Z.clone()
            vs.inducing_directions.data = P.Vz.clone()
            vd.natural_vec.data = nat_vec.clone()
            vd.natural_mat.data = (-0.5 * prec).clone()
            vs.variational_params_initialized.fill_(1)
            model.mean_module.constant.data = P.c.clone()
            model.covar_module.raw_outputscale.data = P.raw_os.clone()
            model.covar_module.base_kernel.raw_lengthscale.data = P.raw_ell.clone()
            likelihood.noise_covar.raw_noise.data = P.raw_noise.clone()
        model.train(), likelihood.train()
        mll = gpytorch.mlls.VariationalELBO(likelihood, model, num_data=num_data)
        opt = gpytorch.optim.NGD(model.variational_parameters(), num_data=num_data, lr=0.1)
        loss = -mll(likelihood(model(x, derivative_directions=Vx)), y)
        loss.backward()
        res = dict(n=n, d=d, M=M, p=p, seed=seed, num_data=num_data, x=x, Vx=Vx, y=y,
                   params={k: v for k, v in P.tensors().items() if k not in ("m", "Ls_raw")},
                   natural_vec=nat_vec, natural_mat=(-0.5 * prec), elbo=(-loss).detach().clone(),
                   grad_natural_vec=-vd.natural_vec.grad.detach().clone(), grad_natural_mat=-vd.natural_mat.grad.detach().clone(),
                   grad_Z=-vs.inducing_points.grad.detach().clone(), grad_raw_noise=-likelihood.noise_covar.raw_noise.grad.detach().clone())
        opt.step()
        res["natural_vec_after"], res["natural_mat_after"] = vd.natural_vec.detach().clone(), vd.natural_mat.detach().clone()
        return res
    finally:
        torch.set_default_dtype(torch.float32)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    f64, f32 = torch.float64, torch.float32
    shared = {"shared_d3_p2_f64": shared_case(30, 3, 8, 2, f64, 31), "shared_d3_p2_f32": shared_case(30, 3, 8, 2, f32, 31),
              "shared_d5_p1_f64": shared_case(25, 5, 12, 1, f64, 32), "shared_d6_p3_f32": shared_case(40, 6, 10, 3, f32, 33)}
    torch.save(shared, os.path.join(out_dir, "shared_cases.pt"))
    print("shared_cases.pt", {k: float(v["elbo"]) for k, v in shared.items()})
    if "--shared-only" in sys.argv:
        return
    kernels = {
        "k_d2_p2_f64": kernel_case(7, 9, 2, 2, f64, 1),
        "k_d3_p1_f64": kernel_case(8, 5, 3, 1, f64, 2),
        "k_d10_p2_f64": kernel_case(6, 11, 10, 2, f64, 3),
        "k_d6_p3_f64_same": kernel_case(9, 9, 6, 3, f64, 4, same=True),
        "k_d10_p2_f32": kernel_case(6, 11, 10, 2, f32, 5),
    }
    torch.save(kernels, os.path.join(out_dir, "kernel_cases.pt"))
    steps = {
        # C1 shape (tests/test_dsvgp.py:21-29) at a reduced minibatch
        "dsvgp_c1_f64": strategy_case("dsvgp", 50, 2, 20, 2, f64, 10),
        "dsvgp_c1_f32": strategy_case("dsvgp", 50, 2, 20, 2, f32, 10),
        "dsvgp_c1_canonical_f64": strategy_case("dsvgp", 30, 2, 20, 2, f64, 11, perturb_dirs=False),
        "dsvgp_d3_p1_f64": strategy_case("dsvgp", 40, 3, 16, 1, f64, 12),
        "dsvgp_d6_p3_f64": strategy_case("dsvgp", 24, 6, 9, 3, f64, 13),
        "dfree_d4_p2_f64": strategy_case("dfree", 30, 4, 10, 2, f64, 14),
        "dfree_d4_p2_f32": strategy_case("dfree", 30, 4, 10, 2, f32, 14),
        "grad_d2_f64": strategy_case("grad", 20, 2, 8, 2, f64, 15),
        "grad_d3_f32": strategy_case("grad", 16, 3, 6, 3, f32, 16),
    }
    torch.save(steps, os.path.join(out_dir, "step_cases.pt"))
    ngd = {"ngd_d3_p1_f64": ngd_case(40, 3, 16, 1, f64, 21), "ngd_c1_f32": ngd_case(50, 2, 20, 2, f32, 22)}
    torch.save(ngd, os.path.join(out_dir, "ngd_cases.pt"))
    print("ngd_cases.pt", {k: float(v["elbo"]) for k, v in ngd.items()})
    for name, blob in (("kernel_cases.pt", kernels), ("step_cases.pt", steps)):
        print(name, {k: (tuple(v["K"].shape) if "K" in v else float(v["elbo"])) for k, v in blob.items()},
              os.path.getsize(os.path.join(out_dir, name)), "bytes")


if __name__ == "__main__":
    main()
