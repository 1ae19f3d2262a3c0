"""In-step kernel timeline of the CUDA-graph training step (TransformerLM_scaled, 64 x 256 tokens) from CUPTI activity
records (torch.profiler): per kernel name the number of launches per step, the mean in-step duration and the idle gap
in front of it, i.e. what the kernels cost INSIDE the replayed graph (warm L2, PDL overlap) as opposed to the
stand-alone probes.  `python tools/step_timeline.py [steps]`   (measurement aid; a profiled run is never a bench value)
"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drakegpt_b200 import model as M  # noqa: E402
from drakegpt_b200.graph import GraphedTrainStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
B, T, V = 64, 256, 80
torch.manual_seed(42)
model = M.TransformerLM(V, 384, T, 6, 6, 0.2, precision="bf16").to(dev).train()
r = model.runner()
r.configure_optimizer(lr=3e-4, betas=(0.9, 0.95))
step = GraphedTrainStep(r, B, T, None)
x, y = torch.randint(0, V, (B, T), device=dev), torch.randint(0, V, (B, T), device=dev)
for _ in range(10):
    step.step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step.step(x, y)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
prev_end = None
span0, span1 = evs[0].time_range.start, evs[-1].time_range.end
# With programmatic dependent launch a kernel STARTS while its predecessor drains (its prologue overlaps, then it waits in
# griddepcontrol.wait), so raw durations overlap and sum to more than the step.  The attributable cost of a kernel is the
# time from its predecessor's end to its own end ("marginal"); marginals sum to the span of the step.
for e in evs:
    a = agg.setdefault(e.name[:96], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
    a[2] += e.time_range.end - (prev_end if prev_end is not None else e.time_range.start)
    prev_end = max(prev_end or 0, e.time_range.end)
print(f"{steps} steps: span {(span1 - span0) / steps:.1f} us/step; raw kernel durations (overlapping) {sum(a[1] for a in agg.values()) / steps:.1f} us/step")
print("  marginal us/step | launches/step | marginal us per launch | raw duration per launch | kernel")
for name, (n, dur, marg) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{marg / steps:8.1f}  n={n / steps:5.1f}  {marg / n:6.2f}  {dur / n:6.2f}  {name}")
