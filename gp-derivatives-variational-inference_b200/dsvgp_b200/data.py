"""Device-resident minibatch sampling + select_cols_of_y in one kernel.

Reference: a shuffling `DataLoader` over a `TensorDataset(x, y)` with y = [f, df/dx_1..d] per row
(directionalvi/directional_vi.py:229-234), then `select_cols_of_y` (:68-90) keeps the value column and
`minibatch_dim` random gradient columns and builds the one-hot directions, which the loop repeats per point (:239)
and interleaves (:241).  Here the dataset is copied to HBM once; an epoch is one device permutation; a minibatch is
ONE launch of `dsvgp_gather_batch_*` that writes x, the interleaved labels and the direction rows.
"""
import ctypes
import random

import torch

from . import _lib


class DeviceMinibatchSampler:
    """sampler = DeviceMinibatchSampler(x, y, batch_size); for idx in sampler.epoch(): x_b, y_b, V = sampler.gather(idx, cols)"""

    def __init__(self, x, y, batch_size, device=None, shuffle=True, generator=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if y.dim() == 1:
            y = y.unsqueeze(-1)
        if x.shape[0] != y.shape[0] or x.dtype != y.dtype:
            raise ValueError("x and y must have the same number of rows and the same dtype")
        self.x = x.to(device).contiguous()
        self.y = y.to(device).contiguous()
        self.batch_size, self.shuffle, self.generator = int(batch_size), shuffle, generator
        self.N, self.d = self.x.shape
        self.ycols = self.y.shape[1]

    def __len__(self):
        return (self.N + self.batch_size - 1) // self.batch_size

    def epoch(self):
        """Yield one int64 device index tensor per minibatch (a fresh permutation per epoch when shuffling)."""
        if self.shuffle:
            perm = torch.randperm(self.N, device=self.x.device, generator=self.generator)
        else:
            perm = torch.arange(self.N, device=self.x.device)
        for s in range(0, self.N, self.batch_size):
            yield perm[s:s + self.batch_size]

    @staticmethod
    def draw_columns(minibatch_dim, dim):
        """The reference's column draw (directional_vi.py:75-77): Python's `random`, value column 0 always kept."""
        return sorted(random.sample(range(1, dim + 1), minibatch_dim) + [0])

    def gather(self, idx, cols, want_directions=True):
        """-> x_batch (n, d), y_batch (n*(p+1),) interleaved, V (n*p, d) one-hot rows (None if p == 0 or not wanted)."""
        n, p = int(idx.shape[0]), len(cols) - 1
        dev, T = self.x.device, self.x.dtype
        xb = torch.empty(n, self.d, dtype=T, device=dev)
        yb = torch.empty(n * (p + 1), dtype=T, device=dev)
        V = torch.empty(n * p, self.d, dtype=T, device=dev) if (want_directions and p) else None
        cols_host = (ctypes.c_int * (p + 1))(*[int(c) for c in cols])
        _lib.call("dsvgp_gather_batch_" + _lib.suffix(T), self.x, self.y, self.N, self.d, self.ycols, idx.contiguous(), n, p,
                  cols_host, xb, yb, V)
        return xb, yb, V
