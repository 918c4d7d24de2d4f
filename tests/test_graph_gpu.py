"""CUDA-graph replay of the training step (dsvgp_b200/graphs.py): same numbers as the eager step, bit for bit, across
minibatches and optimiser updates; the Cholesky status is read after the replay and a failure falls back to the eager
jitter ladder / error path."""
import pytest
import torch

pytestmark = pytest.mark.gpu
F32, F64 = torch.float32, torch.float64


def _setup(wl_over, dtype=F32):
    import bench
    from dsvgp_b200 import gp
    wl = dict(bench.WORKLOADS["C3"], **wl_over)
    dev = torch.device("cuda", 0)
    model, lik = bench.build_model(wl, dtype, dev)
    mll = gp.VariationalELBO(lik, model, num_data=(wl["d"] + 1) * wl["N"])
    batch = lambda seed: tuple(t.to(dev) for t in bench.synth_batch(wl["n"], wl["d"], wl["p"], "dsvgp", dtype, "cpu", seed))
    return wl, model, lik, mll, batch


@pytest.mark.parametrize("M,n,dtype", [(128, 512, F32), (43, 95, F32), (64, 256, F64)])
def test_replayed_step_equals_eager_step(M, n, dtype):
    from dsvgp_b200 import graphs
    wl, model, lik, mll, batch = _setup(dict(M=M, n=n, N=20000), dtype)
    params = list(model.parameters()) + list(lik.parameters())

    def eager(x, V, y):
        for q in params:
            q.grad = None
        loss = -mll(lik(model(x, derivative_directions=V)), y)
        loss.backward()
        return loss.detach().clone(), [q.grad.clone() for q in params]

    ref = [eager(*batch(s)) for s in (1, 2, 3)]
    g = graphs.GraphedStep(model, lik, mll, *batch(1))
    for k, s in enumerate((1, 2, 3, 1)):
        loss = g(*batch(s))
        rl, rg = ref[k % 3]
        assert torch.equal(loss, rl)
        for q, r in zip(params, rg):
            assert torch.equal(q.grad, r)
    assert g.fallbacks == 0


def test_graphed_training_trajectory_equals_eager():
    """20 Adam steps driven by replays == 20 eager steps (parameters move in place underneath the captured graph)."""
    from dsvgp_b200 import graphs
    from dsvgp_b200.optim import FusedAdam

    def run(graphed):
        torch.manual_seed(0)
        wl, model, lik, mll, batch = _setup(dict(M=64, n=512, N=5000))
        params = list(model.parameters()) + list(lik.parameters())
        vd = model.variational_strategy._variational_distribution
        opt = FusedAdam([{"params": params}], lr=0.02, lower_triangular=[vd.chol_variational_covar])
        g = graphs.GraphedStep(model, lik, mll, *batch(100)) if graphed else None
        losses = []
        for it in range(20):
            x, V, y = batch(100 + it)
            if graphed:
                loss = g(x, V, y)
            else:
                opt.zero_grad()
                loss = -mll(lik(model(x, derivative_directions=V)), y)
                loss.backward()
            opt.step()
            losses.append(float(loss))
        return losses, torch.cat([q.detach().reshape(-1) for q in params]).clone()

    le, pe = run(False)
    lg, pg = run(True)
    assert le == lg and torch.equal(pe, pg)
    assert le[-1] < le[0] - 0.5


def test_failed_factorisation_falls_back_to_the_eager_path():
    from dsvgp_b200 import engine, graphs
    wl, model, lik, mll, batch = _setup(dict(M=32, n=128, N=5000), F64)
    g = graphs.GraphedStep(model, lik, mll, *batch(1))
    g()
    with torch.no_grad():
        model.variational_strategy.inducing_points[0, 0] = float("nan")
    with pytest.raises(engine.NanError):
        g()
    assert g.fallbacks == 1
