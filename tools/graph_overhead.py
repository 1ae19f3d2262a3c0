"""How much does a CUDA-graph kernel node cost? Replays a graph of N trivial dependent kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drakegpt_b200 import ops
ctr = torch.zeros(1, device="cuda", dtype=torch.int64)
N = 140
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        ops.raw_counter_add(ctr, 1)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(N):
        ops.raw_counter_add(ctr, 1)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); e1.synchronize()
print(f"graph of {N} trivial kernels: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per replay -> {e0.elapsed_time(e1) / 20 / N * 1e3:.2f} us per node")
