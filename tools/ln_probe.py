"""Time LayerNorm fwd/bwd, colsum, embedding kernels at the scaled-model shape."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drakegpt_b200 import ops
M, C = 16384, 384
dev = "cuda"
x = torch.randn(M, C, device=dev); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
y = torch.empty(M, C, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
dy = torch.randn(M, C, device=dev).bfloat16(); dres = torch.randn(M, C, device=dev); dx = torch.empty(M, C, device=dev)
dxm = torch.empty(M, C, device=dev, dtype=torch.bfloat16); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev); cs = torch.zeros(C, device=dev)
dh = torch.randn(M, 4 * C, device=dev).bfloat16(); csh = torch.zeros(4 * C, device=dev)
idx = torch.randint(0, 80, (64, 256), device=dev); dtok = torch.zeros(80, C, device=dev); dpos = torch.zeros(256, C, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
drop = ops.Dropout(0.2, 1, 3)
ops.raw_ln_fwd(x, g, b, y, mean, rstd)
cases = {
    "ln_fwd": lambda: ops.raw_ln_fwd(x, g, b, y, mean, rstd),
    "ln_bwd(+mask+colsum)": lambda: ops.raw_ln_bwd(dy, x, g, mean, rstd, dres, dx, dg, db, dxm=dxm, dropout=drop, dxm_colsum=cs),
    "colsum[16384x1536 bf16]": lambda: ops.raw_colsum(dh, csh, accumulate=True),
    "embed_bwd": lambda: ops.raw_embed_bwd(idx, dx.view(64, 256, C), dtok, dpos),
}
for name, fn in cases.items():
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name}: {statistics.mean(ts):.1f} us", flush=True)
