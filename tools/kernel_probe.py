"""Device time per launch of the non-GEMM kernels at the scaled-model shape (B=64, T=256, C=384, NH=6).

Like tools/gemm_probe.py: R independent buffer sets (together larger than the 126 MB L2), all launches of a
case captured in one CUDA graph, replays timed with CUDA events -> no host launch latency in the number.
`python tools/kernel_probe.py [case-substring ...]`
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drakegpt_b200 import ops

dev = "cuda"
B, T, NH, H, C, V = 64, 256, 6, 64, 384, 80
D, M, F = NH * H, B * T, 4 * C
R = 6
want = sys.argv[1:]


def attn_set(p):
    qkv = (torch.randn(B, T, 3 * D, device=dev) * 0.5).bfloat16()
    go = torch.randn(B, T, D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    o = torch.empty(B, T, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, NH, T, device=dev)
    dx = torch.empty_like(qkv)
    scr = torch.empty(16, device=dev)
    drop = ops.Dropout(p, 1, 0) if p else None
    fwd = lambda: ops.raw_attn_fwd(q, k, v, o, lse, NH, H, H ** -0.5, drop)
    bwd = lambda: ops.raw_attn_bwd(q, k, v, o, lse, go, dx[:, :, :D], dx[:, :, D:2 * D], dx[:, :, 2 * D:], scr, NH, H,
                                   H ** -0.5, drop)
    fwd()
    return fwd, bwd


def ln_set():
    x = torch.randn(M, C, device=dev); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    y = torch.empty(M, C, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    dy = torch.randn(M, C, device=dev).bfloat16(); dres = torch.randn(M, C, device=dev); dx = torch.empty(M, C, device=dev)
    dxm = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev); cs = torch.zeros(C, device=dev)
    drop = ops.Dropout(0.2, 1, 3)
    fwd = lambda: ops.raw_ln_fwd(x, g, b, y, mean, rstd)
    bwd = lambda: ops.raw_ln_bwd(dy, x, g, mean, rstd, dres, dx, dg, db, dxm=dxm, dropout=drop, dxm_colsum=cs)
    fwd()
    return fwd, bwd


def misc_set():
    dh = torch.randn(M, F, device=dev).bfloat16(); csh = torch.zeros(F, device=dev)
    idx = torch.randint(0, V, (B, T), device=dev)
    tok = torch.randn(V, C, device=dev); pos = torch.randn(T, C, device=dev)
    x = torch.empty(B, T, C, device=dev); dx = torch.randn(B, T, C, device=dev)
    dtok = torch.zeros(V, C, device=dev); dpos = torch.zeros(T, C, device=dev)
    return (lambda: ops.raw_colsum(dh, csh, accumulate=True),
            lambda: ops.raw_embed_fwd(idx, tok, pos, x),
            lambda: ops.raw_embed_bwd(idx, dx, dtok, dpos))


cases = {}
a2, a0, ln, ms = [attn_set(0.2) for _ in range(R)], [attn_set(0.0) for _ in range(R)], [ln_set() for _ in range(R)], [misc_set() for _ in range(R)]
cases["attn_fwd p=0.2"] = [s[0] for s in a2]
cases["attn_bwd p=0.2"] = [s[1] for s in a2]
cases["attn_fwd p=0"] = [s[0] for s in a0]
cases["attn_bwd p=0"] = [s[1] for s in a0]
cases["ln_fwd"] = [s[0] for s in ln]
cases["ln_bwd(+mask+colsum)"] = [s[1] for s in ln]
cases["colsum[16384x1536 bf16]"] = [s[0] for s in ms]
cases["embed_fwd"] = [s[1] for s in ms]
cases["embed_bwd"] = [s[2] for s in ms]

for name, fns in cases.items():
    if want and not any(w in name for w in want):
        continue
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for rep in range(4):
            for f in fns:
                f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); e1.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / (5 * 4 * len(fns)):.1f} us", flush=True)
