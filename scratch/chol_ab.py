"""A/B of the two diagonal-block chains of the blocked Cholesky (csrc/chol.cu): correctness against torch.linalg and time.
    python scratch/chol_ab.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch
from dsvgp_b200 import ops, _lib
F64 = torch.float64
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())
for Mq in (250, 448, 1024, 1500, 3072, 3200):
    g = torch.Generator().manual_seed(Mq)
    R = torch.randn(Mq, Mq + 5, generator=g, dtype=F64)
    A = (R @ R.T / (Mq + 5) + 1e-3 * torch.eye(Mq, dtype=F64)).cuda()
    Lr = torch.linalg.cholesky(A)
    Wr = torch.linalg.inv(Lr)
    Mp, nb0, nlev = ops.chol_plan(Mq)
    for variant in (1, 2, 3):
        _lib.call_raw("dsvgp_set_chol_variant", variant)
        Aw = torch.empty(Mp, Mp, dtype=F64, device="cuda")
        L = torch.full((Mp, Mp), float("nan"), dtype=F64, device="cuda")
        W = torch.full((Mp, Mp), float("nan"), dtype=F64, device="cuda")
        info = torch.ones(1, dtype=torch.int32, device="cuda")

        def run():
            Aw.zero_()
            Aw[:Mq, :Mq] = A
            ops.pad_identity(Aw, Mq)
            ops.cholesky_inverse(Aw, L, W, nb0, nlev, info)
        run()
        torch.cuda.synchronize()
        eL, eW = rel(L[:Mq, :Mq].tril(), Lr), rel(W[:Mq, :Mq].tril(), Wr)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            run()
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        tot = e0.elapsed_time(e1) / 10
        e0.record()
        for _ in range(10):
            Aw.zero_()
            Aw[:Mq, :Mq] = A
            ops.pad_identity(Aw, Mq)
        e1.record()
        torch.cuda.synchronize()
        print(f"Mq={Mq} Mp={Mp} nb0={nb0} nlev={nlev} variant {variant}: info {int(info.item())} errL {eL:.1e} errW {eW:.1e} "
              f"chol+inv {tot - e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
_lib.call_raw("dsvgp_set_chol_variant", 3)
