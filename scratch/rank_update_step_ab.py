"""A/B of the rank-K fp64 kernel modes (dsvgp_set_rank_update 0 / 2 / 1) on whole steps, eager and CUDA-graph replayed.
The same loop with ops.set_chol_inv_streams(1 / 2) in place of set_rank_update gave the one- / two-stream inverse comparison
(equal); the per-level layout is timed by scratch/chol_graph_ab.py."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch, bench
from dsvgp_b200 import ops, graphs
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for name, n in (("C3", 512), ("C3", 4096), ("C2", 500)):
    arm = bench.Arm(dict(bench.WORKLOADS[name]), dev, 0, 1)
    out = []
    for rnd in range(2):
        for ns in (0, 2, 1):
            ops.set_rank_update(ns)
            e = arm.time_steps(n, 1000, 3, 20, collective=False)
            g = graphs.time_graphed_step(arm, n, 2000, 3, 20)
            out.append((ns, round(e, 3), round(g, 3)))
    print(name, n, "(rank_update mode, eager ms, graph ms):", out, flush=True)
    arm.release()
ops.set_rank_update(0)
