"""Full-gradient SVGP drivers (reference directionalvi/grad_svgp.py): inducing variables carry the value and all d
partial derivatives; GradVariationalStrategy on gpytorch's RBFKernelGrad.  Same names / signatures."""
import sys

import torch

from dsvgp_b200 import gp

import directional_vi as _dvi
from utils.count_params import count_params


class GPModel(gp.ApproximateGP):
    """reference grad_svgp.py:20-40."""

    def __init__(self, inducing_points, **kwargs):
        dim = inducing_points.size(1)
        if kwargs.get("variational_strategy") == "CIQ":
            raise NotImplementedError("the CIQ strategy is outside the B200 hot path")
        vd_class = gp.NaturalVariationalDistribution if kwargs.get("variational_distribution") == "NGD" \
            else gp.CholeskyVariationalDistribution
        variational_distribution = vd_class(inducing_points.size(0) * (dim + 1))
        variational_strategy = gp.GradVariationalStrategy(self, inducing_points, variational_distribution,
                                                          learn_inducing_locations=True)
        super().__init__(variational_strategy)
        self.mean_module = gp.ConstantMean()
        self.covar_module = gp.ScaleKernel(gp.RBFKernelGrad())

    def forward(self, x):
        return gp.MultivariateNormal(self.mean_module(x), self.covar_module(x))


def train_gp(train_dataset, dim, num_inducing=128, minibatch_size=1, num_epochs=1, use_ngd=False, use_ciq=False,
             learning_rate_hypers=0.01, learning_rate_ngd=0.1, lr_sched=None, mll_type="ELBO",
             num_contour_quadrature=15, watch_model=False, gamma=0.1, verbose=True, **args):
    """reference grad_svgp.py:42-171.  num_data = N here, not (dim+1)*N (:119, quirk Q4)."""
    if use_ciq:
        raise NotImplementedError("use_ciq (contour-integral-quadrature whitening) is outside the B200 hot path")
    device = _dvi._require_cuda()
    n_samples = len(train_dataset)
    dtype = train_dataset[0][0].dtype
    model = GPModel(torch.rand(num_inducing, dim).to(dtype),
                    **({"variational_distribution": "NGD"} if use_ngd else {})).to(device=device, dtype=dtype)
    likelihood = gp.GaussianLikelihood().to(device=device, dtype=dtype)
    model.train()
    likelihood.train()
    if verbose:
        count_params(model, likelihood)
    vopt, hopt, vsched, hsched = _dvi._optimizers(model, likelihood, learning_rate_hypers, lr_sched, n_samples,
                                                  minibatch_size, num_epochs, gamma,
                                                  ngd=dict(num_data=n_samples, lr=learning_rate_ngd) if use_ngd else None)
    mll_cls = {"ELBO": gp.VariationalELBO, "PLL": gp.PredictiveLogLikelihood}[mll_type]
    mll = mll_cls(likelihood, model, num_data=n_samples)
    total_step, loss = 0, None
    for i in range(num_epochs):
        mini_steps = 0
        for x_batch, y_batch in _dvi._batches(train_dataset, minibatch_size, True, device):
            y_batch = y_batch.reshape(torch.numel(y_batch))
            vopt.zero_grad()
            hopt.zero_grad()
            output = likelihood(model(x_batch))
            loss = -mll(output, y_batch)
            loss.backward()
            vopt.step()
            vsched.step()
            hopt.step()
            hsched.step()
            if total_step % 25 == 0 and verbose:
                means, stds = output.mean[::dim + 1], output.variance.sqrt()[::dim + 1]
                nll = -torch.distributions.Normal(means, stds).log_prob(y_batch[::dim + 1]).mean()
                print(f"Epoch: {i}; total_step: {mini_steps}, loss: {loss.item()}, nll: {nll}")
                sys.stdout.flush()
            mini_steps += 1
            total_step += 1
    if verbose and loss is not None:
        print(f"Done! loss: {loss.item()}")
    print("\nDone Training!")
    return model, likelihood


def eval_gp(test_dataset, model, likelihood, mll_type="ELBO", num_inducing=128, minibatch_size=1):
    """reference grad_svgp.py:173-202."""
    device = _dvi._require_cuda()
    model.eval()
    likelihood.eval()
    means, variances = [], []
    with torch.no_grad():
        for x_batch, _ in _dvi._batches(test_dataset, minibatch_size, False, device):
            preds = likelihood(model(x_batch))
            means.append(preds.mean.cpu())
            variances.append(preds.variance.cpu())
    return torch.cat(means), torch.cat(variances)
