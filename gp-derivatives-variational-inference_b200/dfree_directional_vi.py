"""Derivative-free DSVGP drivers (reference directionalvi/dfree_directional_vi.py): the labels are function values
only; the inducing variables still carry p directional derivatives.  Same names / signatures as the reference."""
import sys

import torch

from dsvgp_b200 import gp

import directional_vi as _dvi
from utils.count_params import count_params


class GPModel(_dvi.GPModel):
    """reference dfree_directional_vi.py:26-60 (always learns the inducing locations)."""

    strategy_class = gp.DFreeDirectionalGradVariationalStrategy

    def __init__(self, inducing_points, inducing_directions, dim, **kwargs):
        super().__init__(inducing_points, inducing_directions, dim, learn_inducing_locations=True, **kwargs)


def train_gp(train_dataset, num_inducing=128, num_directions=1, minibatch_size=1, minibatch_dim=1, num_epochs=1,
             learning_rate_hypers=0.01, learning_rate_ngd=0.1, inducing_data_initialization=True, use_ngd=False,
             use_ciq=False, lr_sched=None, mll_type="ELBO", num_contour_quadrature=15, watch_model=False, gamma=0.1,
             verbose=True, **args):
    """reference dfree_directional_vi.py:93-257.  num_data = (dim+1)*N as in the reference (:130, quirk Q4)."""
    assert num_directions == minibatch_dim
    if use_ciq:
        raise NotImplementedError("use_ciq (contour-integral-quadrature whitening) is outside the B200 hot path")
    device = _dvi._require_cuda()
    dim = len(train_dataset[0][0])
    n_samples = len(train_dataset)
    num_data = (dim + 1) * n_samples
    inducing_points, inducing_directions = _dvi._initial_inducing(train_dataset, num_inducing, num_directions, dim,
                                                                  inducing_data_initialization)
    dtype = train_dataset[0][0].dtype
    model = GPModel(inducing_points.to(dtype), inducing_directions.to(dtype), dim,
                    **({"variational_distribution": "NGD"} if use_ngd else {})).to(device=device, dtype=dtype)
    likelihood = gp.GaussianLikelihood().to(device=device, dtype=dtype)
    model.train()
    likelihood.train()
    if verbose:
        count_params(model, likelihood)
    vopt, hopt, vsched, hsched = _dvi._optimizers(model, likelihood, learning_rate_hypers, lr_sched, n_samples,
                                                  minibatch_size, num_epochs, gamma,
                                                  ngd=dict(num_data=num_data, lr=learning_rate_ngd) if use_ngd else None)
    mll_cls = {"ELBO": gp.VariationalELBO, "PLL": gp.PredictiveLogLikelihood}[mll_type]
    mll = mll_cls(likelihood, model, num_data=num_data)
    total_step, loss = 0, None
    for i in range(num_epochs):
        for x_batch, y_batch in _dvi._batches(train_dataset, minibatch_size, True, device):
            derivative_directions = torch.eye(dim, dtype=dtype)[:num_directions].repeat(len(x_batch), 1)
            vopt.zero_grad()
            hopt.zero_grad()
            output = likelihood(model(x_batch, derivative_directions=derivative_directions))
            loss = -mll(output, y_batch.reshape(-1))
            loss.backward()
            vopt.step()
            vsched.step()
            hopt.step()
            hsched.step()
            if total_step % 50 == 0 and verbose:
                nll = -torch.distributions.Normal(output.mean, output.variance.sqrt()).log_prob(y_batch.reshape(-1)).mean()
                print(f"Epoch: {i}; total_step: {total_step}, loss: {loss.item()}, nll: {nll}")
                sys.stdout.flush()
            total_step += 1
    if verbose and loss is not None:
        print(f"Done! loss: {loss.item()}")
        print("\nDone Training!")
    return model, likelihood


def eval_gp(test_dataset, model, likelihood, mll_type="ELBO", num_directions=1, minibatch_size=1, minibatch_dim=1):
    """reference dfree_directional_vi.py:260-292: one mean / variance per test point."""
    return _dvi.eval_gp(test_dataset, model, likelihood, mll_type, num_directions, minibatch_size, minibatch_dim)
