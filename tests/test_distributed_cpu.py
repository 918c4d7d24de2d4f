"""world_size-2 gloo test (CPU) of the data-parallel cut of the DSVGP step: shard the minibatch,
normalise the data term by the GLOBAL n', sum the per-rank payload ONCE, add the replicated KL term -- the result
must equal the unsharded ELBO and its gradients.  The per-shard arithmetic is done by the CPU oracle (tests may use
it); what is under test is the product's host logic in dsvgp_b200/distributed.py: shard bounds, payload reduction
over a real process group, and the reduce-then-tail = tail-then-reduce claim (linearity)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dsvgp_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data_term(P, x, Vx, y, n_global_outputs):
    mean, var = O.predictive(P, x, Vx)
    s2 = O.noise(P)
    var = O.clamp_variance(var + s2)
    terms = -0.5 * (((y - mean) ** 2 + var) / s2 + torch.log(s2) + 1.8378770664093453)
    return terms.sum() / n_global_outputs


def _worker(rank, world, port, n, d, M, p, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dsvgp_b200 import distributed
        P, x, Vx, y, nd = O.make_problem(n, d, M, p, torch.float64, seed=3)
        lo, hi = distributed.shard_bounds(n, rank, world)
        q = p + 1
        Q = P.clone().requires_grad_(True)
        val = _data_term(Q, x[lo:hi], Vx[lo * p: hi * p], y[lo * q: hi * q], n * q)
        names = list(Q.tensors())
        grads = torch.autograd.grad(val, [getattr(Q, k) for k in names])
        small = torch.cat([val.detach().reshape(1)] + [g.reshape(-1) for g in grads])
        big = torch.zeros(4, dtype=torch.float64)          # stands for [G | t]; only its summation is checked here
        big[rank] = 1.0
        reducer = distributed.Reducer()                    # the two-phase form the engine drives: begin([G|t]) ... end(small)
        reducer.begin(big)
        reducer.end(small)
        assert torch.equal(big[:world], torch.ones(world, dtype=torch.float64))
        chk_b, chk_s = torch.ones(3, dtype=torch.float64), torch.full((2,), float(rank), dtype=torch.float64)
        distributed.all_reduce_payload(chk_b, chk_s)       # blocking form
        assert torch.equal(chk_b, torch.full((3,), float(world), dtype=torch.float64))
        assert torch.equal(chk_s, torch.full((2,), float(sum(range(world))), dtype=torch.float64))
        # replicated tail: the KL term is added once on every rank, never reduced
        Q2 = P.clone().requires_grad_(True)
        kl = O.kl_divergence(Q2) / nd
        gk = torch.autograd.grad(kl, [Q2.m, Q2.Ls_raw])
        elbo = small[0] - kl.detach()
        off, got = 1, {}
        for k in names:
            t = getattr(P, k)
            got[k] = small[off: off + t.numel()].reshape(t.shape).clone()
            off += t.numel()
        got["m"] -= gk[0]
        got["Ls_raw"] -= gk[1]
        ref_val, ref = O.elbo_and_grads(P, x, Vx, y, nd)
        assert abs(float(elbo - ref_val)) < 1e-12 * abs(float(ref_val))
        for k in names:
            assert float((got[k] - ref[k]).abs().max()) <= 1e-11 * float(ref[k].abs().max() + 1e-300), k
        # every rank ends with identical gradients (no second collective needed)
        gathered = [torch.zeros_like(small) for _ in range(world)]
        dist.all_gather(gathered, small)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_step_equals_unsharded_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 37, 3, 6, 2, out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}


def test_shard_bounds_cover_exactly():
    from dsvgp_b200 import distributed
    for n in (1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            spans = [distributed.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_enable_requires_process_group():
    from dsvgp_b200 import distributed
    with pytest.raises(RuntimeError):
        distributed.enable(object(), 10)
