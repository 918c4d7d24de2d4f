"""Shared-direction DSVGP drivers (reference directionalvi/shared_directional_vi.py): one set of `num_directions`
inducing directions for all inducing points and a variational distribution over M + p values.  Same names / signatures
as the reference; the loop is directional_vi's."""
import torch

from dsvgp_b200 import gp

import directional_vi as _dvi
from directional_vi import eval_gp, select_cols_of_y  # noqa: F401  (identical in the reference: :68-90, :271-305)


class GPModel(gp.ApproximateGP):
    """reference shared_directional_vi.py:25-65: `inducing_directions` is the (p, d) shared set; q(u) has M + p values."""

    def __init__(self, inducing_points, inducing_directions, dim, learn_inducing_locations=True, **kwargs):
        self.num_inducing = len(inducing_points)
        self.num_directions = len(inducing_directions)
        if kwargs.get("variational_strategy") == "CIQ":
            raise NotImplementedError("the CIQ strategy is outside the B200 hot path (SURVEY.md section 2, row 7)")
        vd_class = gp.NaturalVariationalDistribution if kwargs.get("variational_distribution") == "NGD" \
            else gp.CholeskyVariationalDistribution
        variational_distribution = vd_class(self.num_inducing + self.num_directions)
        variational_strategy = gp.SharedDirectionalGradVariationalStrategy(
            self, inducing_points, inducing_directions, variational_distribution,
            learn_inducing_locations=learn_inducing_locations)
        super().__init__(variational_strategy)
        self.mean_module = gp.ConstantMean()
        self.covar_module = gp.ScaleKernel(gp.RBFKernelDirectionalGrad())

    def forward(self, x, **params):
        return gp.MultivariateNormal(self.mean_module(x), self.covar_module(x, **params))


def train_gp(train_dataset, num_inducing=128, num_directions=1, minibatch_size=1, minibatch_dim=1, num_epochs=1,
             learning_rate_hypers=0.01, learning_rate_ngd=0.1, inducing_data_initialization=True, use_ngd=False,
             use_ciq=False, lr_sched=None, mll_type="ELBO", num_contour_quadrature=15, watch_model=False, gamma=0.1,
             verbose=True, fixed_inducing_locations=None, **args):
    """reference shared_directional_vi.py:93-268.  The shared direction set starts at the first `num_directions`
    canonical directions (:154); with inducing_data_initialization=True the reference repeats it per inducing point
    (:145), which its own strategy then cannot consume -- here the (p, d) set is used in both cases."""
    return _dvi.train_gp(train_dataset, num_inducing, num_directions, minibatch_size, minibatch_dim, num_epochs,
                         learning_rate_hypers, learning_rate_ngd, inducing_data_initialization, use_ngd, use_ciq, lr_sched,
                         mll_type, num_contour_quadrature, watch_model, gamma, verbose, fixed_inducing_locations,
                         _model_factory=_make_model, **args)


def _make_model(inducing_points, inducing_directions, dim, num_inducing, num_directions, **kw):
    shared = torch.eye(dim, dtype=inducing_points.dtype)[:num_directions]
    return GPModel(inducing_points, shared, dim, **kw)
