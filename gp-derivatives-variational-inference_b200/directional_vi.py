"""DSVGP training / prediction drivers with the reference's names and signatures
(directionalvi/directional_vi.py: GPModel :25, select_cols_of_y :68, train_gp :93, eval_gp :271), running on the
B200 engine.  The minibatch loop is host Python exactly as in the reference; each step is one fused GPU pass.

Differences that are deliberate and visible:
  * a CUDA device is required (no CPU path); tensors of a TensorDataset are moved to the GPU once and minibatches
    are sliced there instead of going through a per-sample DataLoader (other Dataset types still use DataLoader);
  * use_ciq (the contour-integral-quadrature strategy) is outside the hot path -> NotImplementedError; use_ngd
    (NaturalVariationalDistribution + NGD) is supported.
"""
import random
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

from dsvgp_b200 import gp
from dsvgp_b200.data import DeviceMinibatchSampler
from dsvgp_b200.optim import FusedAdam

from utils.count_params import count_params


class GPModel(gp.ApproximateGP):
    """reference directional_vi.py:25-65 -- same constructor, same sub-module names, same state-dict keys."""

    strategy_class = gp.DirectionalGradVariationalStrategy

    def __init__(self, inducing_points, inducing_directions, dim, learn_inducing_locations=True, **kwargs):
        self.num_inducing = len(inducing_points)
        self.num_directions = int(len(inducing_directions) / self.num_inducing)
        if kwargs.get("variational_strategy") == "CIQ":
            raise NotImplementedError("the CIQ strategy is outside the B200 hot path (SURVEY.md section 2, row 7)")
        vd_class = gp.NaturalVariationalDistribution if kwargs.get("variational_distribution") == "NGD" \
            else gp.CholeskyVariationalDistribution                                     # reference :38-43
        variational_distribution = vd_class(self.num_inducing * (self.num_directions + 1))
        variational_strategy = self.strategy_class(self, inducing_points, inducing_directions, variational_distribution,
                                                   learn_inducing_locations=learn_inducing_locations)
        super().__init__(variational_strategy)
        self.mean_module = gp.ConstantMean()
        self.covar_module = gp.ScaleKernel(gp.RBFKernelDirectionalGrad())

    def forward(self, x, **params):
        return gp.MultivariateNormal(self.mean_module(x), self.covar_module(x, **params))


def select_cols_of_y(y_batch, minibatch_dim, dim):
    """Keep the function-value column and `minibatch_dim` randomly chosen gradient columns; return the matching
    canonical directions (reference :68-90; Python's `random` draws the columns there too)."""
    cols = sorted(random.sample(range(1, dim + 1), minibatch_dim) + [0])
    eye = torch.eye(dim, device=y_batch.device)
    return y_batch[:, cols], eye[np.array(cols[1:]) - 1]


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("dsvgp_b200 needs a CUDA device: the hot path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def _batches(dataset, batch_size, shuffle, device):
    """Yield (x, y) minibatches on `device`."""
    if isinstance(dataset, TensorDataset):
        tensors = [t.to(device, non_blocking=True) for t in dataset.tensors]
        n = tensors[0].shape[0]
        perm = torch.randperm(n, device=device) if shuffle else None
        for s in range(0, n, batch_size):
            if perm is None:
                yield tuple(t[s:s + batch_size] for t in tensors)
            else:
                idx = perm[s:s + batch_size]
                yield tuple(t[idx] for t in tensors)
    else:
        for batch in DataLoader(dataset, batch_size=batch_size, shuffle=shuffle):
            yield tuple(t.to(device) for t in batch)


def _train_batches(dataset, batch_size, minibatch_dim, dim, device, dtype):
    """Yield (x, interleaved y of the selected columns, per-point direction rows) for one epoch.  A TensorDataset
    stays resident in HBM and every minibatch is one fused gather launch (dsvgp_b200/data.py); the column draw is
    the reference's (`random.sample`, :75-77), so the same `random.seed` selects the same columns."""
    if isinstance(dataset, TensorDataset) and len(dataset.tensors) == 2 and dataset.tensors[0].dtype == dataset.tensors[1].dtype:
        sampler = getattr(dataset, "_dsvgp_sampler", None)
        if sampler is None or sampler.batch_size != batch_size or sampler.x.device != device:
            sampler = DeviceMinibatchSampler(dataset.tensors[0], dataset.tensors[1], batch_size, device)
            dataset._dsvgp_sampler = sampler
        for idx in sampler.epoch():
            cols = DeviceMinibatchSampler.draw_columns(minibatch_dim, dim)
            yield sampler.gather(idx, cols)
        return
    for x_batch, y_batch in _batches(dataset, batch_size, True, device):
        y_batch, derivative_directions = select_cols_of_y(y_batch, minibatch_dim, dim)
        yield (x_batch, y_batch.reshape(torch.numel(y_batch)),          # interleaved [f, d1..dp] per point
               derivative_directions.to(dtype).repeat(y_batch.size(0), 1))


def _initial_inducing(train_dataset, num_inducing, num_directions, dim, inducing_data_initialization):
    inducing_directions = torch.eye(dim)[:num_directions].repeat(num_inducing, 1)
    if inducing_data_initialization is True:
        inducing_points = torch.stack([train_dataset[i][0] for i in range(num_inducing)]).clone().cpu()
    else:
        inducing_points = torch.rand(num_inducing, dim)
    return inducing_points, inducing_directions


def _optimizers(model, likelihood, lr, lr_sched, n_samples, minibatch_size, num_epochs, gamma, ngd=None):
    # same two optimisers and parameter groups as the reference (:186-199); each FusedAdam.step() is one fused launch
    vd = model.variational_strategy._variational_distribution
    if ngd is not None:                       # use_ngd: natural-gradient steps on the natural parameters (:187)
        variational_optimizer = gp.NGD(model.variational_parameters(), num_data=ngd["num_data"], lr=ngd["lr"])
    else:
        variational_optimizer = FusedAdam([{"params": model.variational_parameters()}], lr=lr,
                                          lower_triangular=[vd.chol_variational_covar])
    hyperparameter_optimizer = FusedAdam([{"params": model.hyperparameters()},
                                          {"params": likelihood.parameters()}], lr=lr)
    if lr_sched == "step_lr":
        num_batches = int(np.ceil(n_samples / minibatch_size))
        milestones = [int(num_epochs * num_batches / 3), int(2 * num_epochs * num_batches / 3)]
        mk = lambda o: torch.optim.lr_scheduler.MultiStepLR(o, milestones, gamma=gamma)
    else:
        fn = (lambda epoch: 1.0) if lr_sched is None else lr_sched
        mk = lambda o: torch.optim.lr_scheduler.LambdaLR(o, lr_lambda=fn)
    return variational_optimizer, hyperparameter_optimizer, mk(variational_optimizer), mk(hyperparameter_optimizer)


def train_gp(train_dataset, num_inducing=128, num_directions=1, minibatch_size=1, minibatch_dim=1, num_epochs=1,
             learning_rate_hypers=0.01, learning_rate_ngd=0.1, inducing_data_initialization=True, use_ngd=False,
             use_ciq=False, lr_sched=None, mll_type="ELBO", num_contour_quadrature=15, watch_model=False, gamma=0.1,
             verbose=True, fixed_inducing_locations=None, _model_factory=None, **args):
    """Train a DSVGP (reference :93-268).  Returns (model, likelihood).
    (_model_factory: the sibling drivers that differ only in the model they build -- shared_directional_vi -- plug in here.)"""
    assert num_directions == minibatch_dim
    if use_ciq:
        raise NotImplementedError("use_ciq (contour-integral-quadrature whitening) is outside the B200 hot path")
    device = _require_cuda()
    dim = len(train_dataset[0][0])
    n_samples = len(train_dataset)
    num_data = (dim + 1) * n_samples

    inducing_points, inducing_directions = _initial_inducing(train_dataset, num_inducing, num_directions, dim,
                                                             inducing_data_initialization)
    learn_inducing_locations = True
    if fixed_inducing_locations is not None:
        inducing_points, learn_inducing_locations = fixed_inducing_locations, False
    dtype = train_dataset[0][0].dtype
    mkw = dict(learn_inducing_locations=learn_inducing_locations, **({"variational_distribution": "NGD"} if use_ngd else {}))
    if _model_factory is not None:
        model = _model_factory(inducing_points.to(dtype), inducing_directions.to(dtype), dim, num_inducing, num_directions, **mkw)
    else:
        model = GPModel(inducing_points.to(dtype), inducing_directions.to(dtype), dim, **mkw)
    model = model.to(device=device, dtype=dtype)
    likelihood = gp.GaussianLikelihood().to(device=device, dtype=dtype)
    model.train()
    likelihood.train()
    if verbose:
        count_params(model, likelihood)

    variational_optimizer, hyperparameter_optimizer, variational_scheduler, hyperparameter_scheduler = _optimizers(
        model, likelihood, learning_rate_hypers, lr_sched, n_samples, minibatch_size, num_epochs, gamma,
        ngd=dict(num_data=num_data, lr=learning_rate_ngd) if use_ngd else None)
    if mll_type == "ELBO":
        mll = gp.VariationalELBO(likelihood, model, num_data=num_data)
    elif mll_type == "PLL":
        mll = gp.PredictiveLogLikelihood(likelihood, model, num_data=num_data)
    else:
        raise ValueError(f"unknown mll_type {mll_type!r}")

    total_step = 0
    loss = None
    for i in range(num_epochs):
        for x_batch, y_batch, derivative_directions in _train_batches(train_dataset, minibatch_size, minibatch_dim, dim,
                                                                      device, dtype):
            kwargs = {"derivative_directions": derivative_directions}
            variational_optimizer.zero_grad()
            hyperparameter_optimizer.zero_grad()
            output = likelihood(model(x_batch, **kwargs))
            loss = -mll(output, y_batch)
            loss.backward()
            variational_optimizer.step()
            variational_scheduler.step()
            hyperparameter_optimizer.step()
            hyperparameter_scheduler.step()
            if total_step % 50 == 0 and verbose:
                means = output.mean[::num_directions + 1]
                stds = output.variance.sqrt()[::num_directions + 1]
                nll = -torch.distributions.Normal(means, stds).log_prob(y_batch[::num_directions + 1]).mean()
                print(f"Epoch: {i}; total_step: {total_step}, loss: {loss.item()}, nll: {nll}")
                sys.stdout.flush()
            total_step += 1
    if verbose and loss is not None:
        print(f"Done! loss: {loss.item()}")
        print("\nDone Training!")
    return model, likelihood


def eval_gp(test_dataset, model, likelihood, mll_type="ELBO", num_directions=1, minibatch_size=1, minibatch_dim=1):
    """Predict means and variances (incl. likelihood noise) for every output of every test point, concatenated on
    the CPU (reference :271-305)."""
    assert num_directions == minibatch_dim
    device = _require_cuda()
    dim = len(test_dataset[0][0])
    model.eval()
    likelihood.eval()
    means, variances = [], []
    with torch.no_grad():
        for x_batch, _ in _batches(test_dataset, minibatch_size, False, device):
            derivative_directions = torch.eye(dim, dtype=x_batch.dtype)[:num_directions].repeat(len(x_batch), 1)
            preds = likelihood(model(x_batch, derivative_directions=derivative_directions))
            means.append(preds.mean.cpu())
            variances.append(preds.variance.cpu())
    print("Done Testing!")
    return torch.cat(means), torch.cat(variances)
