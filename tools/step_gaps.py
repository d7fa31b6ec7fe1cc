"""Where the CIRIM step's time goes at B = 16: CUDA-event time of one forward vs the sum of its kernel times (torch.profiler,
CUPTI), kernels grouped by name."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mridc_b200 as mb
from mridc_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda")
torch.manual_seed(1)
model = mb.CIRIM(synth.cirim_cfg("GRU")).eval().to(dev)
d = {k: v.to(dev) for k, v in synth.make_batch(B, 15, 320, 320).items() if isinstance(v, torch.Tensor)}
fwd = lambda: next(model(d["y"], d["sensitivity_maps"], d["mask"], None, d["target"]))[-1][-1]
for _ in range(3): fwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fwd(); fwd(); e1.record(); torch.cuda.synchronize()
print("forward: %.2f ms (CUDA events, mean of 2)" % (e0.elapsed_time(e1) / 2))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fwd(); torch.cuda.synchronize()
agg = collections.OrderedDict()
t0, t1 = None, None
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg.setdefault(ev.name[:60], [0, 0.0]); a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        s, e = ev.time_range.start, ev.time_range.end
        t0 = s if t0 is None else min(t0, s); t1 = e if t1 is None else max(t1, e)
tot = sum(a[1] for a in agg.values())
print("kernel time %.2f ms in %d launches; first-to-last span %.2f ms" % (tot / 1e3, sum(a[0] for a in agg.values()), (t1 - t0) / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("  %-60s n=%4d %9.1f us %5.1f %%" % (k, a[0], a[1], 100 * a[1] / tot))
