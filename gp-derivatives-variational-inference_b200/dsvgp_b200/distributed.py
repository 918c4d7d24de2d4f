"""Data-parallel DSVGP step: one process per GPU, the minibatch sharded over ranks, ONE exchange per step.

The reference has no distributed code; this is the only parallelism the hot path needs
(section 8e).  Rank r takes a contiguous shard of the minibatch.  Everything of size proportional to n (K_zx, A, B, C,
mean, variance, residuals, and their backward up to the weighted Gram matrix G = A diag(g) A^T) is local.  What is
summed over ranks, once per step, are two buffers the engine lays out contiguously for exactly this purpose:

    big   = [ G (M'^2) | t (M') ]                         model dtype
    small = [ scalars (8) | dZ (M d) | dV_z (M p d) ]     float64   (data term, d noise, d c, and the K_zx-backward
                                                                     contributions to Z, V_z, lengthscale, outputscale)

after which every rank runs the identical O(M'^3) tail (Cholesky backward is linear in its upstream gradient, so
reduce-then-tail equals tail-then-reduce) and ends with identical parameter gradients -- no second collective.
The data term is normalised by the GLOBAL n' (the reference divides by the batch's event size, VariationalELBO).
"""
import os

import torch
import torch.distributed as dist



def shard_bounds(n, rank, world_size):
    """Contiguous, balanced [lo, hi) of a length-n minibatch for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_payload(big, small, group=None):
    """Sum the two engine buffers over `group` (blocking form).  `big` may be None (forward-only evaluation)."""
    if big is not None:
        dist.all_reduce(big, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(small, op=dist.ReduceOp.SUM, group=group)


class Reducer:
    """The one exchange of a sharded step, split in two so that it hides under compute:

        begin(big)   right after the Gram product: the all-reduce of [G | t] (37.8 MB at C3) is issued asynchronously --
                     NCCL runs it on its own stream -- while the main stream goes on with dK_zx = L^-T dA and the K_zx
                     assembly backward, neither of which needs G;
        end(small)   after those: the small fp64 buffer they accumulated into follows, and the main stream then waits for
                     both collectives before the replicated tail.

    One Reducer per model (model.variational_strategy._reducer), so that other models of the process stay unsharded."""

    def __init__(self, group=None, shard_tail=True, overlap=None):
        self.group = group
        # overlap=False: the [G | t] all-reduce is issued blocking at `end` (after dK_zx / kdir_bwd) instead of underneath them
        self.overlap = (os.environ.get("DSVGP_REDUCE_OVERLAP", "1") != "0") if overlap is None else bool(overlap)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.shard_tail = shard_tail      # also split the replicated O(M'^3) Cholesky-backward tail over the ranks (engine._tail_panel)
        self._pending = None

    def begin(self, big):
        if self.overlap:
            self._pending = dist.all_reduce(big, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            self._big = big

    def end(self, small):
        if not self.overlap and getattr(self, "_big", None) is not None:
            dist.all_reduce(self._big, op=dist.ReduceOp.SUM, group=self.group)
            self._big = None
        w = dist.all_reduce(small, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        if self._pending is not None:
            self._pending.wait()
            self._pending = None
        w.wait()


    def reduce_tail(self, small2):
        """second, tiny exchange of a step with a sharded tail: every rank's share of dK_zz contracted with dK_zz/dtheta"""
        dist.all_reduce(small2, op=dist.ReduceOp.SUM, group=self.group)


def enable(model, n_global, group=None, shard_tail=True):
    """Make every subsequent ELBO step of `model` a shard of a global minibatch of `n_global` points."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    vs = model.variational_strategy
    vs._n_global = int(n_global)
    vs._reducer = Reducer(group, shard_tail)


def disable(model):
    vs = model.variational_strategy
    vs._n_global = None
    vs._reducer = None


def broadcast_parameters(model, likelihood, src=0, group=None):
    """Replicas must start from identical parameters (they stay identical because gradients are)."""
    with torch.no_grad():
        for t in list(model.state_dict().values()) + list(likelihood.state_dict().values()):
            dist.broadcast(t, src=src, group=group)
