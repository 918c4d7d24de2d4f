"""Fused optimiser step for the DSVGP training loop.

The reference steps two `torch.optim.Adam` instances after every ELBO forward+backward
(directionalvi/directional_vi.py:192-199, :251-254): one over `model.variational_parameters()` and one over
`model.hyperparameters()` + `likelihood.parameters()`, each followed by a per-minibatch LR scheduler.  `FusedAdam`
is a drop-in `torch.optim.Optimizer` with the same constructor, `param_groups`, `state` layout (`step`, `exp_avg`,
`exp_avg_sq`) and therefore the same `state_dict()`, so `MultiStepLR` / `LambdaLR` and checkpoints work unchanged --
but `step()` is ONE launch of `dsvgp_adam_step_*` over every tensor of the optimiser instead of torch's ~10
foreach kernels per dtype.  There is no CPU path: parameters must live on a CUDA device.
"""
import ctypes

import torch

from . import _lib

_MAX_GROUPS = 4


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics (amsgrad / maximize / capturable are not
    offered), one kernel launch per dtype per step.

    `lower_triangular`: parameters (identity-compared) that are square matrices whose strictly-upper gradient is
    structurally zero -- `chol_variational_covar`; only their lower triangle is read and written."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, lower_triangular=()):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tri = {id(p) for p in lower_triangular}
        if len(self.param_groups) > _MAX_GROUPS:
            raise ValueError(f"FusedAdam supports at most {_MAX_GROUPS} parameter groups")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        per_dtype = {}
        ghost = (ctypes.c_double * (8 * len(self.param_groups)))()
        for gi, group in enumerate(self.param_groups):
            step_t = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _lib.DsvgpError("FusedAdam: parameters must be CUDA tensors (there is no CPU path)")
                if p.grad.is_sparse or not p.is_contiguous():
                    raise _lib.DsvgpError("FusedAdam: dense contiguous parameters only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)                    # host counter, like torch's non-capturable Adam
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step_t = float(st["step"])
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                # (with weight decay the strictly-upper entries are not fixed points of the update: visit everything)
                tri = p.shape[0] if (id(p) in self._tri and p.dim() == 2 and p.shape[0] == p.shape[1]
                                     and group["weight_decay"] == 0) else 0
                per_dtype.setdefault(p.dtype, []).append((p, g, st["exp_avg"], st["exp_avg_sq"], gi, tri, step_t))
            b1, b2 = group["betas"]
            ghost[8 * gi: 8 * gi + 5] = [float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                         float(group["weight_decay"])]
        for dtype, items in per_dtype.items():
            # one bias-correction step per group: tensors that joined a group later (grad was None before) get their
            # own launch so that every tensor sees its own step count, as in torch
            by_step = {}
            for it in items:
                by_step.setdefault((it[4], it[6]), []).append(it)
            steps_of_group = {}
            for (gi, stp) in by_step:
                steps_of_group.setdefault(gi, set()).add(stp)
            launches = [items] if all(len(v) == 1 for v in steps_of_group.values()) else list(by_step.values())
            for batch in launches:
                for it in batch:
                    ghost[8 * it[4] + 5] = it[6]
                desc = (ctypes.c_int64 * (8 * len(batch)))()
                for k, (p, g, m, v, gi, tri, _) in enumerate(batch):
                    desc[8 * k: 8 * k + 8] = [p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), gi, tri, 0]
                _lib.call("dsvgp_adam_step_" + _lib.suffix(dtype), len(batch), desc, len(self.param_groups), ghost)
        return loss
