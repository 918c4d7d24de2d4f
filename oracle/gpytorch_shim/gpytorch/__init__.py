"""Stand-in for gpytorch==1.4.0 -- TEST INFRASTRUCTURE ONLY, never imported by the product.

gpytorch is pinned by the reference (graphite_environment.yml:96) but is not installed in
this image and there is no network.  This single file re-states, from the published 1.4.0
behaviour, exactly the pieces the reference's hot-path files import, so that the UNMODIFIED
reference files under /root/reference/directionalvi can be executed here on CPU by
`oracle/make_golden.py` to produce the golden vectors in tests/golden/.  Everything is dense
and eager (no lazy evaluation, no CG, no batching): the numerical results are the same, only
the laziness is gone.

Each piece says which gpytorch 1.4.0 component it stands in for.  Nothing in here is reference
code; nothing in here is shipped.
"""
import math
import sys
import types
import warnings

import torch
from torch.nn import Parameter
from torch.nn.functional import softplus

__version__ = "1.4.0-shim"


def _submodule(name):
    full = __name__ + "." + name
    mod = types.ModuleType(full)
    sys.modules[full] = mod
    parent, _, leaf = full.rpartition(".")
    setattr(sys.modules[parent], leaf, mod)
    return mod


# --------------------------------------------------------------------------- settings
settings = _submodule("settings")


class _Flag:
    _state = False

    @classmethod
    def on(cls):
        return cls._state

    @classmethod
    def off(cls):
        return not cls._state


class trace_mode(_Flag):
    pass


class debug(_Flag):
    pass


class _Value:
    _global_value = None

    def __init__(self, value=None):
        self._new = value

    @classmethod
    def value(cls, *args):
        return cls._global_value

    @classmethod
    def _set_value(cls, v):
        cls._global_value = v

    def __enter__(self):
        self._old = type(self)._global_value
        type(self)._global_value = self._new

    def __exit__(self, *a):
        type(self)._global_value = self._old


class cholesky_jitter(_Value):
    """gpytorch.settings.cholesky_jitter: 1e-6 (float) / 1e-8 (double); .value() with no dtype
    returns the float value (the reference calls it that way, DGVS.py:74)."""

    @classmethod
    def value(cls, dtype=None):
        if dtype is not None and (dtype == torch.float64 or getattr(dtype, "dtype", None) == torch.float64):
            return 1e-8
        return 1e-6


class min_variance(_Value):
    @classmethod
    def value(cls, dtype=None):
        if torch.is_tensor(dtype):
            dtype = dtype.dtype
        return 1e-10 if dtype == torch.float64 else 1e-6


class num_contour_quadrature(_Value):
    _global_value = 15


class max_preconditioner_size(_Value):
    _global_value = 15


class max_cg_iterations(_Value):
    _global_value = 1000


class eval_cg_tolerance(_Value):
    _global_value = 1e-2


class cg_tolerance(_Value):
    _global_value = 1.0


class max_lanczos_quadrature_iterations(_Value):
    _global_value = 20


for _c in (trace_mode, debug, cholesky_jitter, min_variance, num_contour_quadrature, max_preconditioner_size,
           max_cg_iterations, eval_cg_tolerance, cg_tolerance, max_lanczos_quadrature_iterations):
    setattr(settings, _c.__name__, _c)

# --------------------------------------------------------------------------- utils
utils = _submodule("utils")
errors = _submodule("utils.errors")
warnings_mod = _submodule("utils.warnings")
cholesky_mod = _submodule("utils.cholesky")
memoize = _submodule("utils.memoize")
broadcasting = _submodule("utils.broadcasting")


class NanError(RuntimeError):
    pass


class NotPSDError(RuntimeError):
    pass


class CachingError(RuntimeError):
    pass


class OldVersionWarning(UserWarning):
    pass


class NumericalWarning(RuntimeWarning):
    pass


errors.NanError, errors.NotPSDError, errors.CachingError = NanError, NotPSDError, CachingError
warnings_mod.OldVersionWarning, warnings_mod.NumericalWarning = OldVersionWarning, NumericalWarning


def psd_safe_cholesky(A, upper=False, out=None, jitter=None, max_tries=3):
    """gpytorch.utils.cholesky.psd_safe_cholesky: plain Cholesky, on failure a NaN check and the
    jitter ladder jitter*10**i, i < max_tries, then NotPSDError."""
    L, info = torch.linalg.cholesky_ex(A)
    if not bool(info.any()):
        return L.transpose(-1, -2) if upper else L
    if torch.isnan(A).any():
        raise NanError(f"cholesky_cpu: {int(torch.isnan(A).sum())} of {A.numel()} elements are NaN.")
    if jitter is None:
        jitter = cholesky_jitter.value(A.dtype)
    Aprime = A.clone()
    jitter_prev = 0
    for i in range(max_tries):
        jitter_new = jitter * (10 ** i)
        Aprime.diagonal(dim1=-2, dim2=-1).add_(jitter_new - jitter_prev)
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not bool(info.any()):
            warnings.warn(f"A not p.d., added jitter of {jitter_new:.1e} to the diagonal", NumericalWarning)
            return L.transpose(-1, -2) if upper else L
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter_new:.1e}.")


cholesky_mod.psd_safe_cholesky = psd_safe_cholesky


def cached(method=None, name=None, ignore_args=False):
    """gpytorch.utils.memoize.cached: per-object memo in obj._memoize_cache."""
    def deco(fn):
        key = name if name is not None else fn.__name__

        def wrapper(self, *args, **kwargs):
            cache = self.__dict__.setdefault("_memoize_cache", {})
            k = key if ignore_args or not (args or kwargs) else (key, args, tuple(sorted(kwargs.items())))
            if k not in cache:
                cache[k] = fn(self, *args, **kwargs)
            return cache[k]
        wrapper.__name__ = fn.__name__
        return wrapper
    return deco(method) if method is not None else deco


def clear_cache_hook(module, *a, **k):
    module.__dict__["_memoize_cache"] = {}


def pop_from_cache_ignore_args(obj, name):
    try:
        return obj.__dict__["_memoize_cache"].pop(name)
    except KeyError:
        raise CachingError(f"{name} not in cache")


memoize.cached, memoize.clear_cache_hook, memoize.pop_from_cache_ignore_args = cached, clear_cache_hook, pop_from_cache_ignore_args
broadcasting._mul_broadcast_shape = lambda *shapes: torch.broadcast_shapes(*shapes)
utils.linear_cg = None  # only the (out-of-scope) CIQ strategy uses it

# --------------------------------------------------------------------------- lazy tensors (dense, eager)
lazy = _submodule("lazy")
_kron = _submodule("lazy.kronecker_product_lazy_tensor")


class LazyTensor:
    """Dense eager stand-in for gpytorch.lazy.LazyTensor and its subclasses."""

    def __init__(self, t):
        self._t = t.evaluate() if isinstance(t, LazyTensor) else t

    def evaluate(self):
        return self._t

    @property
    def shape(self):
        return self._t.shape

    @property
    def dtype(self):
        return self._t.dtype

    @property
    def device(self):
        return self._t.device

    def size(self, *a):
        return self._t.size(*a)

    def dim(self):
        return self._t.dim()

    def diag(self):
        return self._t.diagonal(dim1=-2, dim2=-1)

    def add_jitter(self, jitter_val=1e-3):
        # LazyTensor.add_jitter default is 1e-3 (gpytorch/lazy/lazy_tensor.py)
        n = self._t.shape[-1]
        return LazyTensor(self._t + jitter_val * torch.eye(n, dtype=self._t.dtype, device=self._t.device))

    def add_diag(self, diag):
        n = self._t.shape[-1]
        return LazyTensor(self._t + diag * torch.eye(n, dtype=self._t.dtype, device=self._t.device))

    def mul(self, c):
        return LazyTensor(self._t * c)

    def matmul(self, other):
        o = other.evaluate() if isinstance(other, LazyTensor) else other
        return self._t @ o

    __matmul__ = matmul

    def __add__(self, other):
        o = other.evaluate() if isinstance(other, LazyTensor) else other
        return LazyTensor(self._t + o)

    def transpose(self, a, b):
        return LazyTensor(self._t.transpose(a, b))

    def double(self):
        return LazyTensor(self._t.double())

    def to(self, *a, **k):
        return LazyTensor(self._t.to(*a, **k))

    def __getitem__(self, idx):
        return LazyTensor(self._t[idx])

    def cholesky(self, upper=False):
        return TriangularLazyTensor(psd_safe_cholesky(self._t, upper=upper))

    def root_decomposition(self):
        return RootLazyTensor(psd_safe_cholesky(self._t))

    def logdet(self):
        return torch.logdet(self._t)


class NonLazyTensor(LazyTensor):
    pass


class DiagLazyTensor(LazyTensor):
    def __init__(self, diag):
        self._diag = diag
        super().__init__(torch.diag_embed(diag))

    def cholesky(self, upper=False):
        return DiagLazyTensor(self._diag.sqrt())

    def mul(self, c):
        return DiagLazyTensor(self._diag * c)


class TriangularLazyTensor(LazyTensor):
    def __init__(self, t, upper=False):
        super().__init__(t)
        self.upper = upper

    def inv_matmul(self, rhs):
        return torch.linalg.solve_triangular(self._t, rhs, upper=self.upper)


class RootLazyTensor(LazyTensor):
    def __init__(self, root):
        root = root.evaluate() if isinstance(root, LazyTensor) else root
        self.root = LazyTensor(root)
        super().__init__(root @ root.transpose(-1, -2))


class CholLazyTensor(RootLazyTensor):
    """gpytorch.lazy.CholLazyTensor: covariance L L^T kept with its factor; matmul goes through
    the factor (L (L^T rhs)) and logdet is sum(log(diag(L)^2))."""

    def __init__(self, chol):
        chol = chol.evaluate() if isinstance(chol, LazyTensor) else chol
        super().__init__(chol)
        self._chol = chol

    def matmul(self, other):
        o = other.evaluate() if isinstance(other, LazyTensor) else other
        return self._chol @ (self._chol.transpose(-1, -2) @ o)

    __matmul__ = matmul

    def logdet(self):
        return self._chol.diagonal(dim1=-2, dim2=-1).pow(2).log().sum(-1)


class SumLazyTensor(LazyTensor):
    def __init__(self, *parts):
        self.parts = [p if isinstance(p, LazyTensor) else LazyTensor(p) for p in parts]
        self._full = None

    def evaluate(self):
        if self._full is None:
            self._full = sum(p.evaluate() for p in self.parts)
        return self._full

    @property
    def _t(self):
        return self.evaluate()

    def matmul(self, other):
        return sum(p.matmul(other) for p in self.parts)

    __matmul__ = matmul

    def diag(self):
        return sum(p.diag() for p in self.parts)


class MatmulLazyTensor(LazyTensor):
    """diag() is (left * right^T).sum(-1), as gpytorch's MatmulLazyTensor.diag does."""

    def __init__(self, left, right):
        self.left = left.evaluate() if isinstance(left, LazyTensor) else left
        self.right = right.evaluate() if isinstance(right, LazyTensor) else right
        self._full = None

    def evaluate(self):
        if self._full is None:
            self._full = self.left @ self.right
        return self._full

    @property
    def _t(self):
        return self.evaluate()

    def diag(self):
        return (self.left * self.right.transpose(-1, -2)).sum(-1)


class KroneckerProductLazyTensor(LazyTensor):
    pass


def lazify(x):
    return x if isinstance(x, LazyTensor) else LazyTensor(x)


def delazify(x):
    return x.evaluate() if isinstance(x, LazyTensor) else x


for _c in (LazyTensor, NonLazyTensor, DiagLazyTensor, TriangularLazyTensor, RootLazyTensor, CholLazyTensor,
           SumLazyTensor, MatmulLazyTensor):
    setattr(lazy, _c.__name__, _c)
lazy.lazify, lazy.delazify = lazify, delazify
_kron.KroneckerProductLazyTensor = KroneckerProductLazyTensor

# --------------------------------------------------------------------------- distributions
distributions = _submodule("distributions")


class MultivariateNormal:
    """gpytorch.distributions.MultivariateNormal: mean + (lazy) covariance; .variance is the
    covariance diagonal clamped below at settings.min_variance."""

    def __init__(self, mean, covariance_matrix, validate_args=False):
        self.loc = mean
        self._covar = covariance_matrix if isinstance(covariance_matrix, LazyTensor) else LazyTensor(covariance_matrix)

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return self._covar.evaluate()

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    @property
    def variance(self):
        var = self._covar.diag()
        mv = min_variance.value(var.dtype)
        if var.lt(mv).any():
            var = var.clamp_min(mv)
        return var

    @property
    def stddev(self):
        return self.variance.sqrt()


class Delta:
    pass


distributions.MultivariateNormal, distributions.Delta = MultivariateNormal, Delta


def _kl_mvn_mvn(p, q):
    """gpytorch's registered KL(p || q) for two MultivariateNormals, specialised exactly as the
    whitened strategy uses it: q = N(0, I).  0.5*(logdet q - logdet p + tr(q^-1 p) + maha - k)."""
    mean_diffs = q.loc - p.loc
    pc = p.lazy_covariance_matrix
    logdet_p = pc.logdet()
    qc = q.lazy_covariance_matrix.evaluate()
    logdet_q = torch.logdet(qc)
    root_p = pc.root.evaluate() if isinstance(pc, RootLazyTensor) else psd_safe_cholesky(pc.evaluate())
    rhs = torch.cat([mean_diffs.unsqueeze(-1), root_p], -1)
    sol = torch.linalg.solve(qc, rhs)
    trace_plus_inv_quad = (rhs * sol).sum()
    return 0.5 * (logdet_q - logdet_p + trace_plus_inv_quad - float(mean_diffs.size(-1)))


# --------------------------------------------------------------------------- module / constraints
module = _submodule("module")


class Module(torch.nn.Module):
    def __call__(self, *inputs, **kwargs):
        return self.forward(*inputs, **kwargs)

    def _clear_cache(self):
        self.__dict__["_memoize_cache"] = {}

    def register_parameter(self, name, parameter, prior=None):
        # gpytorch.Module.register_parameter names its second argument `parameter`
        torch.nn.Module.register_parameter(self, name, parameter)

    def hyperparameters(self):
        for _, p in self.named_hyperparameters():
            yield p

    def variational_parameters(self):
        for _, p in self.named_variational_parameters():
            yield p

    def named_hyperparameters(self):
        for prefix, mod in self.named_modules():
            if not isinstance(mod, _VariationalDistribution):
                for elem in mod.named_parameters(prefix=prefix, recurse=False):
                    yield elem

    def named_variational_parameters(self):
        for prefix, mod in self.named_modules():
            if isinstance(mod, _VariationalDistribution):
                for elem in mod.named_parameters(prefix=prefix, recurse=False):
                    yield elem


module.Module = Module

# --------------------------------------------------------------------------- means / kernels
means = _submodule("means")
kernels = _submodule("kernels")
_rbf = _submodule("kernels.rbf_kernel")


class ConstantMean(Module):
    """gpytorch.means.ConstantMean: one learned scalar broadcast over all inputs."""

    def __init__(self):
        super().__init__()
        self.register_parameter("constant", Parameter(torch.zeros(1)))

    def forward(self, input):
        return self.constant.expand(input.shape[:-1])


means.ConstantMean = ConstantMean


def postprocess_rbf(dist_mat):
    return dist_mat.div_(-2).exp_()


class Kernel(Module):
    """gpytorch.kernels.Kernel: __call__(x1, x2=None, diag=False, **params) -> (lazy) forward."""
    has_lengthscale = False

    def __init__(self):
        super().__init__()
        if self.has_lengthscale:
            self.register_parameter("raw_lengthscale", Parameter(torch.zeros(1, 1)))

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_lengthscale.dtype).expand(1, 1)
        self.raw_lengthscale.data.copy_(value + torch.log(-torch.expm1(-value)))

    def num_outputs_per_input(self, x1, x2):
        return 1

    def covar_dist(self, x1, x2, diag=False, last_dim_is_batch=False, square_dist=False,
                   dist_postprocess_func=None, postprocess=True, **params):
        """gpytorch Kernel.covar_dist / Distance._sq_dist: mean-centred expanded squared distance,
        diagonal forced to zero only when x1 == x2 and neither requires grad, clamp at 0."""
        assert square_dist and not diag and not last_dim_is_batch
        x1_eq_x2 = torch.equal(x1, x2)
        adjustment = x1.mean(-2, keepdim=True)
        x1 = x1 - adjustment
        x2 = x2 - adjustment
        x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
        x1_pad = torch.ones_like(x1_norm)
        same = x1_eq_x2 and not x1.requires_grad and not x2.requires_grad
        if same:
            x2_norm, x2_pad = x1_norm, x1_pad
        else:
            x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
            x2_pad = torch.ones_like(x2_norm)
        x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
        x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
        res = x1_.matmul(x2_.transpose(-2, -1))
        if same:
            res.diagonal(dim1=-2, dim2=-1).fill_(0)
        res.clamp_min_(0)
        return dist_postprocess_func(res) if (postprocess and dist_postprocess_func is not None) else res

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        x1_ = x1.unsqueeze(-1) if x1.dim() == 1 else x1
        x2_ = x1_ if x2 is None else (x2.unsqueeze(-1) if x2.dim() == 1 else x2)
        res = self.forward(x1_, x2_, diag=diag, **params)
        return res if diag else lazify(res)


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        if diag:
            return torch.ones(x1.shape[:-1], dtype=x1.dtype, device=x1.device)
        x1_ = x1.div(self.lengthscale)
        x2_ = x2.div(self.lengthscale)
        return self.covar_dist(x1_, x2_, square_dist=True, dist_postprocess_func=postprocess_rbf, **params)


class RBFKernelGrad(RBFKernel):
    """gpytorch.kernels.RBFKernelGrad (1.4.0) restated: value / gradient blocks built
    block-contiguously then perfect-shuffled to the interleaved (MultiTask) ordering."""

    def forward(self, x1, x2, diag=False, **params):
        n1, d = x1.shape[-2:]
        n2 = x2.shape[-2]
        assert not diag
        K = torch.zeros(n1 * (d + 1), n2 * (d + 1), device=x1.device, dtype=x1.dtype)
        x1_ = x1.div(self.lengthscale)
        x2_ = x2.div(self.lengthscale)
        outer = x1_.view(n1, 1, d) - x2_.view(1, n2, d)
        outer = outer / self.lengthscale.unsqueeze(-2)
        outer = torch.transpose(outer, -1, -2).contiguous()
        diff = self.covar_dist(x1_, x2_, square_dist=True, dist_postprocess_func=postprocess_rbf, **params)
        K_11 = diff
        K[:n1, :n2] = K_11
        outer1 = outer.view(n1, n2 * d)
        K[:n1, n2:] = outer1 * K_11.repeat(1, d)
        outer2 = outer.transpose(-1, -3).reshape(n2, n1 * d).transpose(-1, -2)
        K[n1:, :n2] = -outer2 * K_11.repeat(d, 1)
        outer3 = outer1.repeat(d, 1) * outer2.repeat(1, d)
        kp = torch.kron(torch.eye(d, d, device=x1.device, dtype=x1.dtype) / self.lengthscale.pow(2),
                        torch.ones(n1, n2, device=x1.device, dtype=x1.dtype))
        chain_rule = kp - outer3
        K[n1:, n2:] = chain_rule * K_11.repeat(d, d)
        pi1 = torch.arange(n1 * (d + 1)).view(d + 1, n1).t().reshape(n1 * (d + 1))
        pi2 = torch.arange(n2 * (d + 1)).view(d + 1, n2).t().reshape(n2 * (d + 1))
        return K[pi1, :][:, pi2]

    def num_outputs_per_input(self, x1, x2):
        return x1.size(-1) + 1


class ScaleKernel(Kernel):
    """gpytorch.kernels.ScaleKernel: outputscale = softplus(raw_outputscale) times the base kernel."""

    def __init__(self, base_kernel):
        super().__init__()
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", Parameter(torch.tensor(0.0)))

    @property
    def outputscale(self):
        return softplus(self.raw_outputscale)

    def forward(self, x1, x2, diag=False, **params):
        orig = self.base_kernel.forward(x1, x2, diag=diag, **params)
        return delazify(orig) * self.outputscale

    def num_outputs_per_input(self, x1, x2):
        return self.base_kernel.num_outputs_per_input(x1, x2)


_rbf.RBFKernel, _rbf.postprocess_rbf = RBFKernel, postprocess_rbf
kernels.Kernel, kernels.RBFKernel, kernels.RBFKernelGrad, kernels.ScaleKernel = Kernel, RBFKernel, RBFKernelGrad, ScaleKernel

# --------------------------------------------------------------------------- variational
variational = _submodule("variational")
_vs = _submodule("variational._variational_strategy")
_nvd = _submodule("variational.natural_variational_distribution")


class _VariationalDistribution(Module):
    pass


class CholeskyVariationalDistribution(_VariationalDistribution):
    """gpytorch.variational.CholeskyVariationalDistribution: q(u) = N(m, L L^T), L = tril(param)."""

    def __init__(self, num_inducing_points, batch_shape=torch.Size([]), mean_init_std=1e-3, **kwargs):
        super().__init__()
        self.num_inducing_points = num_inducing_points
        self.mean_init_std = mean_init_std
        self.register_parameter("variational_mean", Parameter(torch.zeros(num_inducing_points)))
        self.register_parameter("chol_variational_covar", Parameter(torch.eye(num_inducing_points)))

    @property
    def dtype(self):
        return self.variational_mean.dtype

    @property
    def device(self):
        return self.variational_mean.device

    def shape(self):
        return torch.Size([self.num_inducing_points])

    def forward(self):
        chol = self.chol_variational_covar
        lower_mask = torch.ones(chol.shape[-2:], dtype=chol.dtype, device=chol.device).tril(0)
        return MultivariateNormal(self.variational_mean, CholLazyTensor(chol.mul(lower_mask)))

    def initialize_variational_distribution(self, prior_dist):
        self.variational_mean.data.copy_(prior_dist.mean)
        self.variational_mean.data.add_(torch.randn_like(prior_dist.mean), alpha=self.mean_init_std)
        self.chol_variational_covar.data.copy_(prior_dist.lazy_covariance_matrix.cholesky().evaluate())


def _phi_for_cholesky_(A):
    """Modifies A to be the phi function used in differentiating through Cholesky (lower triangle, halved diagonal)."""
    A.tril_().diagonal(offset=0, dim1=-2, dim2=-1).mul_(0.5)
    return A


def _cholesky_backward(dout_dL, L, L_inverse):
    # gpytorch/variational/natural_variational_distribution.py (1.4.0), after torch's cholesky_backward
    A = L.transpose(-1, -2) @ dout_dL
    phi = _phi_for_cholesky_(A)
    grad_input = (L_inverse.transpose(-1, -2) @ phi) @ L_inverse
    return grad_input.add(grad_input.transpose(-1, -2)).mul_(0.5)


class _NaturalToMuVarSqrt(torch.autograd.Function):
    """gpytorch 1.4.0 _NaturalToMuVarSqrt: natural parameters -> (mean, Cholesky factor of the covariance); the
    backward returns the gradient with respect to the EXPECTATION parameters, i.e. the natural gradient."""

    @staticmethod
    def _forward(nat_mean, nat_covar):
        L_inv = psd_safe_cholesky(-2.0 * nat_covar, upper=False)
        eye = torch.eye(L_inv.size(-1), dtype=L_inv.dtype, device=L_inv.device)
        L = torch.linalg.solve_triangular(L_inv, eye, upper=False)
        S = L.transpose(-1, -2) @ L
        mu = (S @ nat_mean.unsqueeze(-1)).squeeze(-1)
        return mu, psd_safe_cholesky(S, upper=False)

    @staticmethod
    def forward(ctx, nat_mean, nat_covar):
        mu, L = _NaturalToMuVarSqrt._forward(nat_mean, nat_covar)
        ctx.save_for_backward(mu, L)
        return mu, L

    @staticmethod
    def backward(ctx, dout_dmu, dout_dL):
        mu, L = ctx.saved_tensors
        eye = torch.eye(L.size(-1), dtype=L.dtype, device=L.device)
        C = torch.linalg.solve_triangular(L, eye, upper=False)
        dout_dSigma = _cholesky_backward(dout_dL, L, C)
        dout_deta1 = dout_dmu - 2 * (dout_dSigma @ mu.unsqueeze(-1)).squeeze(-1)
        return dout_deta1, dout_dSigma


class NaturalVariationalDistribution(_VariationalDistribution):
    """gpytorch.variational.NaturalVariationalDistribution (1.4.0): parameters natural_vec = S^-1 m and
    natural_mat = -1/2 S^-1; used with gpytorch.optim.NGD (directional_vi.py:38-40, :187)."""

    def __init__(self, num_inducing_points, batch_shape=torch.Size([]), mean_init_std=1e-3, **kwargs):
        super().__init__()
        self.num_inducing_points = num_inducing_points
        self.mean_init_std = mean_init_std
        self.register_parameter("natural_vec", Parameter(torch.zeros(num_inducing_points)))
        self.register_parameter("natural_mat", Parameter(torch.eye(num_inducing_points).mul(-0.5)))

    @property
    def dtype(self):
        return self.natural_vec.dtype

    @property
    def device(self):
        return self.natural_vec.device

    def shape(self):
        return torch.Size([self.num_inducing_points])

    def forward(self):
        mean, chol_covar = _NaturalToMuVarSqrt.apply(self.natural_vec, self.natural_mat)
        return MultivariateNormal(mean, CholLazyTensor(chol_covar))

    def initialize_variational_distribution(self, prior_dist):
        prior_prec = prior_dist.covariance_matrix.inverse()
        prior_mean = prior_dist.mean
        noise = torch.randn_like(prior_mean).mul_(self.mean_init_std)
        self.natural_vec.data.copy_((prior_prec @ prior_mean).add_(noise))
        self.natural_mat.data.copy_(prior_prec.mul(-0.5))


class _VariationalStrategy(Module):
    """gpytorch.variational._variational_strategy._VariationalStrategy (1.4.0)."""

    def __init__(self, model, inducing_points, variational_distribution, learn_inducing_locations=True):
        super().__init__()
        object.__setattr__(self, "model", model)
        inducing_points = inducing_points.clone()
        if inducing_points.dim() == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        if learn_inducing_locations:
            self.register_parameter("inducing_points", Parameter(inducing_points))
        else:
            self.register_buffer("inducing_points", inducing_points)
        self._variational_distribution = variational_distribution
        self.register_buffer("variational_params_initialized", torch.tensor(0))

    @property
    def variational_distribution(self):
        return self._variational_distribution()

    def kl_divergence(self):
        return _kl_mvn_mvn(self.variational_distribution, self.prior_distribution)

    def train(self, mode=True):
        if (self.training and not mode) or mode:
            self._clear_cache()
        return super().train(mode=mode)

    def __call__(self, x, prior=False, **kwargs):
        if prior:
            return self.model.forward(x, **kwargs)
        if self.training:
            self._clear_cache()
        if not self.variational_params_initialized.item():
            prior_dist = self.prior_distribution
            self._variational_distribution.initialize_variational_distribution(prior_dist)
            self.variational_params_initialized.fill_(1)
        inducing_points = self.inducing_points
        variational_dist_u = self.variational_distribution
        return self.forward(x, inducing_points, inducing_values=variational_dist_u.mean,
                            variational_inducing_covar=variational_dist_u.lazy_covariance_matrix, **kwargs)


variational._VariationalDistribution = _VariationalDistribution
variational.CholeskyVariationalDistribution = CholeskyVariationalDistribution
variational.NaturalVariationalDistribution = NaturalVariationalDistribution
variational._VariationalStrategy = _VariationalStrategy
_vs._VariationalStrategy = _VariationalStrategy
_nvd.NaturalVariationalDistribution = NaturalVariationalDistribution

# --------------------------------------------------------------------------- models / likelihoods / mlls
models = _submodule("models")
likelihoods = _submodule("likelihoods")
mlls = _submodule("mlls")
optim = _submodule("optim")


class ApproximateGP(Module):
    def __init__(self, variational_strategy):
        super().__init__()
        self.variational_strategy = variational_strategy

    def forward(self, x):
        raise NotImplementedError

    def __call__(self, inputs, prior=False, **kwargs):
        if inputs.dim() == 1:
            inputs = inputs.unsqueeze(-1)
        return self.variational_strategy(inputs, prior=prior, **kwargs)


models.ApproximateGP = ApproximateGP


class _HomoskedasticNoise(Module):
    def __init__(self):
        super().__init__()
        self.register_parameter("raw_noise", Parameter(torch.zeros(1)))

    @property
    def noise(self):
        # GreaterThan(1e-4) constraint: softplus(raw) + lower bound
        return softplus(self.raw_noise) + 1e-4


class GaussianLikelihood(Module):
    """gpytorch.likelihoods.GaussianLikelihood with the default GreaterThan(1e-4) noise constraint."""

    def __init__(self):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise()

    @property
    def noise(self):
        return self.noise_covar.noise

    def __call__(self, input, *args, **kwargs):
        if isinstance(input, MultivariateNormal):
            mean, covar = input.mean, input.lazy_covariance_matrix
            noise = self.noise.expand(mean.shape[-1])
            return input.__class__(mean, SumLazyTensor(covar, DiagLazyTensor(noise)))
        raise NotImplementedError

    def expected_log_prob(self, target, input, *params, **kwargs):
        mean, variance = input.mean, input.variance
        noise = self.noise.expand(mean.shape)
        res = ((target - mean) ** 2 + variance) / noise + noise.log() + math.log(2 * math.pi)
        return res.mul(-0.5)

    def log_marginal(self, observations, function_dist, *params, **kwargs):
        marginal = self(function_dist)
        var = marginal.variance
        return -0.5 * ((observations - marginal.mean) ** 2 / var + var.log() + math.log(2 * math.pi))


likelihoods.GaussianLikelihood = GaussianLikelihood


class _ApproximateMLL(Module):
    def __init__(self, likelihood, model, num_data, beta=1.0, combine_terms=True):
        super().__init__()
        self.likelihood, self.model = likelihood, model
        self.num_data, self.beta, self.combine_terms = num_data, beta, combine_terms

    def forward(self, approximate_dist_f, target, **kwargs):
        num_batch = approximate_dist_f.event_shape[0]
        log_likelihood = self._log_likelihood_term(approximate_dist_f, target, **kwargs).div(num_batch)
        kl_divergence = self.model.variational_strategy.kl_divergence().div(self.num_data / self.beta)
        return log_likelihood - kl_divergence


class VariationalELBO(_ApproximateMLL):
    def _log_likelihood_term(self, variational_dist_f, target, **kwargs):
        return self.likelihood.expected_log_prob(target, variational_dist_f, **kwargs).sum(-1)


class PredictiveLogLikelihood(_ApproximateMLL):
    def _log_likelihood_term(self, approximate_dist_f, target, **kwargs):
        return self.likelihood.log_marginal(target, approximate_dist_f, **kwargs).sum(-1)


mlls.VariationalELBO, mlls.PredictiveLogLikelihood = VariationalELBO, PredictiveLogLikelihood


class NGD(torch.optim.Optimizer):
    """gpytorch.optim.NGD (1.4.0): natural-gradient step  p <- p - lr * num_data * p.grad  (the gradients of a
    NaturalVariationalDistribution are already natural gradients of the per-datum objective)."""

    def __init__(self, params, num_data, lr=0.1):
        self.num_data = num_data
        super().__init__(params, defaults=dict(lr=lr))

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                p.add_(p.grad, alpha=(-group["lr"] * self.num_data))
        return None


optim.NGD = NGD
