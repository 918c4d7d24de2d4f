"""2-GPU NCCL test of the sharded step through the real engine: gradients of the sharded ELBO (one all-reduce of
[G|t] and the small buffer, replicated tail) equal the single-GPU full-minibatch result and are identical on both
ranks.  Skipped when fewer than 2 GPUs are visible (the 1-GPU round-end run)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import bench
        from dsvgp_b200 import distributed, gp
        wl = dict(bench.WORKLOADS["C3"], M=256, n=1024)
        dev = torch.device("cuda", rank)
        n, d, p = wl["n"], wl["d"], wl["p"]
        x, V, y = bench.synth_batch(n, d, p, "dsvgp", torch.float32, "cpu", 5)

        def grads(xs, Vs, ys, sharded):
            model, lik = bench.build_model(wl, torch.float32, dev)
            if sharded:
                distributed.enable(model, n)
            mll = gp.VariationalELBO(lik, model, num_data=(d + 1) * wl["N"])
            loss = -mll(lik(model(xs.to(dev), derivative_directions=Vs)), ys.to(dev))
            loss.backward()
            distributed.disable(model)
            g = torch.cat([q.grad.reshape(-1).double() for q in list(model.parameters()) + list(lik.parameters())])
            return float(loss), g

        lo, hi = distributed.shard_bounds(n, rank, world)
        q = p + 1
        l_s, g_s = grads(x[lo:hi], V[lo * p: hi * p], y[lo * q: hi * q], True)
        l_f, g_f = grads(x, V, y, False)
        assert abs(l_s - l_f) < 1e-5 * abs(l_f), (l_s, l_f)
        assert float((g_s - g_f).abs().max() / g_f.abs().max()) < 1e-4
        both = [torch.zeros_like(g_s) for _ in range(world)]
        dist.all_gather(both, g_s)
        assert torch.equal(both[0], both[1])
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_step_two_gpus():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("dtype,M,n", [(torch.float64, 96, 300), (torch.float32, 256, 1024), (torch.float32, 1024, 512)])
def test_sharded_tail_equals_replicated_tail_on_one_gpu(world, dtype, M, n):
    """The N > 1 form of the Cholesky-backward tail (column panels of dK_zz per rank, contraction summed over ranks) run for
    every rank of a `world` in ONE process: the summed contributions must equal the replicated tail's gradients."""
    import bench
    from dsvgp_b200 import engine, gp
    wl = dict(bench.WORKLOADS["C3"], M=M, n=n, N=50000)
    dev = torch.device("cuda", 0)
    model, lik = bench.build_model(wl, dtype, dev)
    mll = gp.VariationalELBO(lik, model, num_data=(wl["d"] + 1) * wl["N"])
    x, V, y = (t.to(dev) for t in bench.synth_batch(n, wl["d"], wl["p"], "dsvgp", dtype, "cpu", 11))
    params = list(model.parameters()) + list(lik.parameters())

    def grads():
        for q in params:
            q.grad = None
        loss = -mll(lik(model(x, derivative_directions=V)), y)
        loss.backward()
        return float(loss), torch.cat([q.grad.reshape(-1).double() for q in params])

    l0, g0 = grads()
    engine.ENGINE.tail_shards_debug = world
    try:
        l1, g1 = grads()
    finally:
        engine.ENGINE.tail_shards_debug = None
    assert l0 == l1
    err = float((g1 - g0).abs().max() / g0.abs().max())
    assert err < (1e-11 if dtype == torch.float64 else 2e-6), err
    # the panels partition the columns
    Mq, q = M * 3, 3
    cover = sorted(c for r in range(world) for c in engine.Engine.tail_panels(Mq, q, r, world))
    assert cover[0][0] == 0 and cover[-1][1] == Mq and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
