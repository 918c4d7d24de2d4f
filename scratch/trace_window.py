"""Every kernel of one traced C3 step whose start lies in [t_lo, t_hi) ms (start, duration, stream, name)."""
import sys, os
t_lo, t_hi = float(sys.argv[1]), float(sys.argv[2])
sys.argv = ["x", "C3"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trace_step.py")).read().split("evs = [e for e")[0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
starts = [e for e in evs if "hyp_from_raw" in e.name]
t0, t1 = starts[-2].time_range.start, starts[-1].time_range.start
for e in evs:
    s = (e.time_range.start - t0) / 1000
    if t0 <= e.time_range.start < t1 and t_lo <= s < t_hi:
        print(f"{s:8.3f} {(e.time_range.end - e.time_range.start):7.1f}us s{getattr(e, 'device_resource_id', -1)} {e.name.split('(')[0][-48:]}")
