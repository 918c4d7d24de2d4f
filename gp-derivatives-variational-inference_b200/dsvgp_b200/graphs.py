"""CUDA-graph replay of the whole training step for small minibatches.

At the reference's shipped minibatch (n = 512: experiments/synthetic1/run_exp.py:25, tests/test_dsvgp.py:28) one step is
~160 kernel launches of a few microseconds each around a latency-bound Cholesky: the host cannot enqueue them as fast as
the GPU retires them.  `GraphedStep` captures

    loss = -mll(likelihood(model(x, derivative_directions=V)), y);  loss.backward()

ONCE -- both side streams (the engine's assembly stream and the Cholesky's panel/update stream fork and join with events,
which capture follows) and the autograd backward included -- and replays it per minibatch: inputs are copied into static
buffers, parameters are updated in place by the optimiser, gradients land in static `.grad` tensors.

The one thing a graph cannot contain is the host's look at the Cholesky status (DGVS.py:74 psd_safe_cholesky syncs on
cuSOLVER's info too): the status stays in a device int, is read after the replay, and if the un-jittered factorisation
failed the step is re-run eagerly through the 1e-6 / 1e-5 / 1e-4 ladder -- the re-capture-free form of the retry.
"""
import torch

from .engine import ENGINE

MAX_GRAPH_N = 4096          # above this the step is GPU-bound and a graph buys nothing


class GraphedStep:
    def __init__(self, model, likelihood, mll, x, V, y, warmup=3):
        self.model, self.likelihood, self.mll = model, likelihood, mll
        self.params = [q for q in list(model.parameters()) + list(likelihood.parameters()) if q.requires_grad]
        self.x, self.V, self.y = x.clone(), (None if V is None else V.to(x.device).clone()), y.clone()
        dev = x.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):              # eager warm-up off the default stream (allocates every workspace)
            for _ in range(max(1, warmup)):
                self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        for q in self.params:
            q.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._forward_backward().detach()
        st = model.variational_strategy
        n, d = self.x.shape
        self._factor = ENGINE.factor(dev, self.x.dtype, d, st.inducing_points.shape[0], st._p())
        self.fallbacks = 0

    def _kwargs(self):
        return {} if self.V is None else {"derivative_directions": self.V}

    def _forward_backward(self):
        loss = -self.mll(self.likelihood(self.model(self.x, **self._kwargs())), self.y)
        loss.backward()
        return loss

    def _eager(self):
        for q in self.params:
            q.grad = None
        return self._forward_backward()

    def __call__(self, x=None, V=None, y=None, check=True):
        """Replay on a new minibatch of the captured shape (None = keep the static contents).  Returns the loss (a static
        0-dim tensor, overwritten by the next replay); gradients are in the parameters' .grad."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if V is not None and self.V is not None:
            self.V.copy_(V, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        if check and int(self._factor.info.item()) != 0:     # the step's one host synchronisation, as in the eager path
            self.fallbacks += 1
            grads = [q.grad for q in self.params]
            loss = self._eager()                             # eager: takes the jitter ladder / raises NanError, NotPSDError
            for q, g in zip(self.params, grads):             # keep the static .grad tensors the graph writes into
                if g is not None and q.grad is not None and q.grad is not g:
                    g.copy_(q.grad)
                    q.grad = g
            with torch.no_grad():
                self.loss.copy_(loss.detach())
        return self.loss


def time_graphed_step(arm, n, seed, warmup, steps):
    """bench.py helper: ms per replayed step (status read included) at per-GPU minibatch n, single GPU."""
    x, V, y = (t.to(arm.device) for t in arm.batch(n, seed))
    g = GraphedStep(arm.model, arm.lik, arm.mll, x, V, y)
    for _ in range(warmup):
        g()
    ms = arm.timed(lambda: g(), steps, collective=False)
    for q in arm.params:
        q.grad = None
    del g
    return ms
