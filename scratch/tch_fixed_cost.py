"""One-wave experiment: per-tile fixed cost F and per-k-block time t of the 3xFP16 product (T = F + nk * t)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch
from dsvgp_b200 import ops
F16 = torch.float16
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
M, N = 3072, 1536          # 24 x 6 = 144 CTAs = 72 pairs: one wave on 148 SMs
inv = torch.tensor([1.0], device="cuda")
for K in (256, 768, 1536, 3072, 6144):
    A = [torch.randn(M, K, device="cuda").half() for _ in range(2)]
    B = [torch.randn(K, N, device="cuda").half() for _ in range(2)]
    C = torch.empty(M, N, device="cuda")
    Ch = tuple(torch.empty(M, N, device="cuda", dtype=F16) for _ in range(2))
    one = t(lambda: ops.gemm_tch(A, B, C, M, N, K, inv, chunk=1))
    three = t(lambda: ops.gemm_tch(A, B, C, M, N, K, inv, chunk=1, Ch=Ch, c_scale=inv))
    print(f"K={K:5d} nk={K//64:3d}: one output stream {one:7.1f} us, three streams {three:7.1f} us", flush=True)
