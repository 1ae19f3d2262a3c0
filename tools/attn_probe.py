"""Time the tcgen05 attention kernels at the scaled-model shape (B=64, T=256, NH=6, H=64)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drakegpt_b200 import ops
B, T, NH, H = 64, 256, 6, 64
D = NH * H
qkv = (torch.randn(B, T, 3 * D, device="cuda") * 0.5).bfloat16()
go = torch.randn(B, T, D, device="cuda").bfloat16()
q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
o = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, NH, T, device="cuda")
dx = torch.empty_like(qkv); scr = torch.empty(16, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for p in (0.2, 0.0):
    drop = ops.Dropout(p, 1, 0) if p else None
    for name, fn in (("fwd", lambda: ops.raw_attn_fwd(q, k, v, o, lse, NH, H, H ** -0.5, drop)),
                     ("bwd", lambda: ops.raw_attn_bwd(q, k, v, o, lse, go, dx[:, :, :D], dx[:, :, D:2 * D], dx[:, :, 2 * D:], scr, NH, H, H ** -0.5, drop))):
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"attn {name} p={p}: {statistics.mean(ts):.1f} us", flush=True)
