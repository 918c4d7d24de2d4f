"""Kernel timeline of ONE sharded training step on rank 0 (torch.profiler), launched with torchrun on N GPUs:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29700 scratch/trace_step_dist.py"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-derivatives-variational-inference_b200")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist, bench
from torch.profiler import profile, ProfilerActivity
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
wl = dict(bench.WORKLOADS["C3"])
arm = bench.Arm(wl, dev, rank, world)
arm.shard(wl["n"] * world)
x, V, y = (t.to(dev) for t in arm.batch(wl["n"], 1000 + rank))
for _ in range(4): arm.step(x, V, y)
arm.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    arm.step(x, V, y); torch.cuda.synchronize()
arm.barrier()
if rank == 0:
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    print(f"world {world}: {len(ev)} kernels, span {(ev[-1].time_range.end - t0) / 1e3:.3f} ms")
    # everything after the Gram product's split-K reduce
    start = max(i for i, e in enumerate(ev) if "splitk_reduce" in e.name)
    for e in ev[start:]:
        print(f"  {(e.time_range.start - t0) / 1e3:8.3f} ms  {e.time_range.end - e.time_range.start:8.1f} us  {e.name[:70]}")
dist.destroy_process_group()
