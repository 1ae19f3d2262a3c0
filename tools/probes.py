"""Device-time probes of the training step's kernels at the TransformerLM_scaled shape (B=64, T=256, C=384, NH=6).

Shared by bench.py (the `roofline` object and the per-kernel-family table) and the tools/*_probe.py scripts.
Every case is launched on R independent operand / output sets (together larger than the 126 MB L2, so operands come
from HBM as they do inside the training step), all launches of a case are captured in ONE CUDA graph, and the graph
replay is timed with CUDA events on the launching stream: device time per launch, free of host launch latency.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drakegpt_b200 import ops  # noqa: E402
from drakegpt_b200._lib import MAJOR_K, MAJOR_MN  # noqa: E402

B, T, NH, H, C, V = 64, 256, 6, 64, 384, 80
D, M, F = NH * H, B * T, 4 * C
R = 6

GEMM_SHAPES = {  # name: (M, N, K, a_major, b_major, out dtype, epilogue) -- exactly as engine.py launches them
    "qkv_fwd": (M, 3 * D, C, MAJOR_K, MAJOR_K, torch.bfloat16, "none"),
    "proj_fwd": (M, C, D, MAJOR_K, MAJOR_K, torch.float32, "bias_drop_res"),
    "ffn1_fwd": (M, F, C, MAJOR_K, MAJOR_K, torch.bfloat16, "bias_relu"),
    "ffn2_fwd": (M, C, F, MAJOR_K, MAJOR_K, torch.float32, "bias_drop_res"),
    "ffn2_dgrad": (M, F, C, MAJOR_K, MAJOR_MN, torch.bfloat16, "relu_mask"),
    "ffn1_dgrad": (M, C, F, MAJOR_K, MAJOR_MN, torch.bfloat16, "none"),
    "qkv_dgrad": (M, C, 3 * D, MAJOR_K, MAJOR_MN, torch.bfloat16, "none"),
    "proj_dgrad": (M, D, C, MAJOR_K, MAJOR_MN, torch.bfloat16, "none"),
    "ffn1_wgrad": (F, C, M, MAJOR_MN, MAJOR_MN, torch.float32, "splitk_cs"),
    "ffn2_wgrad": (C, F, M, MAJOR_MN, MAJOR_MN, torch.float32, "splitk_cs"),
    "qkv_wgrad": (3 * D, C, M, MAJOR_MN, MAJOR_MN, torch.float32, "splitk"),
    "proj_wgrad": (C, D, M, MAJOR_MN, MAJOR_MN, torch.float32, "splitk_cs"),
}


def _splits(rows, cols, k):
    from drakegpt_b200.engine import Runner
    return Runner._splits(rows, cols, k, ops.sm_count())


def make_gemm(name, dev="cuda"):
    m, n, k, am, bm, odt, epi = GEMM_SHAPES[name]
    A = torch.randn((m, k) if am == MAJOR_K else (k, m), device=dev).bfloat16()
    Bm = torch.randn((n, k) if bm == MAJOR_K else (k, n), device=dev).bfloat16()
    out = torch.zeros(m, n, device=dev, dtype=odt)
    kw = {}
    if epi == "bias_relu":
        kw = dict(bias=torch.zeros(n, device=dev), relu=True,
                  relu_mask_out=torch.zeros((n // 32) * m, device=dev, dtype=torch.int32))
    elif epi == "bias_drop_res":
        kw = dict(bias=torch.zeros(n, device=dev), residual=torch.zeros(m, n, device=dev), dropout=ops.Dropout(0.2, 1, 1))
    elif epi == "relu_mask":
        kw = dict(relu_mask_in=torch.randint(-2 ** 31, 2 ** 31 - 1, ((n // 32) * m,), device=dev, dtype=torch.int32))
    elif epi == "splitk":
        kw = dict(accumulate=True, split_k=_splits(m, n, k))
    elif epi == "splitk_cs":
        from drakegpt_b200.engine import Runner
        kw = dict(accumulate=True, split_k=Runner._splits(m, n, k, ops.sm_count(), colsum=True), a_colsum=torch.zeros(m, device=dev))
    return lambda: ops.raw_gemm(A, Bm, out, a_major=am, b_major=bm, **kw)


def make_gemm_ln(name, dev="cuda"):
    """The residual GEMM `name` ("proj_fwd" / "ffn2_fwd") with the following LayerNorm in its epilogue."""
    m, n, k = GEMM_SHAPES[name][:3]
    A = torch.randn(m, k, device=dev).bfloat16()
    W = torch.randn(n, k, device=dev).bfloat16()
    res, out = torch.randn(m, n, device=dev), torch.empty(m, n, device=dev)
    bias, g, b = torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)
    y = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(m, device=dev), torch.empty(m, device=dev)
    drop = ops.Dropout(0.2, 1, 1)
    return lambda: ops.raw_gemm_res_ln(A, W, bias, res, out, g, b, y, mean, rstd, dropout=drop)


def gemm_ln_bytes(name):
    m, n, k = GEMM_SHAPES[name][:3]
    return 2.0 * (m * k + n * k) + m * n * (4 + 4 + 2.0) + 8.0 * m  # operands, residual in, x_out + bf16 y out, stats


def gemm_flops(name):
    m, n, k = GEMM_SHAPES[name][:3]
    return 2.0 * m * n * k


def gemm_bytes(name):
    """Algorithmic HBM bytes of one launch: both bf16 operands once + the output once (+ residual / mask)."""
    m, n, k, _, _, odt, epi = GEMM_SHAPES[name]
    b = 2.0 * (m * k + n * k) + m * n * (2 if odt == torch.bfloat16 else 4)
    if epi == "bias_drop_res":
        b += 4.0 * m * n
    if epi in ("bias_relu", "relu_mask"):
        b += m * n / 8.0
    return b


def attn_set(p, dev="cuda"):
    qkv = (torch.randn(B, T, 3 * D, device=dev) * 0.5).bfloat16()
    go = torch.randn(B, T, D, device=dev).bfloat16()
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    o = torch.empty(B, T, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, NH, T, device=dev)
    dx = torch.empty_like(qkv)
    scr = torch.empty(16, device=dev)
    drop = ops.Dropout(p, 1, 0) if p else None

    def fwd():
        ops.raw_attn_fwd(q, k, v, o, lse, NH, H, H ** -0.5, drop)

    def bwd():
        ops.raw_attn_bwd(q, k, v, o, lse, go, dx[:, :, :D], dx[:, :, D:2 * D], dx[:, :, 2 * D:], scr, NH, H, H ** -0.5, drop)

    fwd()
    return fwd, bwd


ATTN_FLOPS_FWD = 4.0 * B * NH * H * T * (T + 1) / 2          # QK^T + PV over the causal pairs
ATTN_FLOPS_BWD = 10.0 * B * NH * H * T * (T + 1) / 2         # S, dP, dV, dK, dQ
ATTN_BYTES_FWD = 2.0 * M * D * 4 + 4.0 * B * NH * T          # q, k, v in; o out (bf16); lse
ATTN_BYTES_BWD = 2.0 * M * D * 8 + 4.0 * B * NH * T          # q, k, v, o, dO in; dq, dk, dv out


def ln_set(dev="cuda"):
    x = torch.randn(M, C, device=dev)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    y = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    dy = torch.randn(M, C, device=dev).bfloat16()
    dres, dx = torch.randn(M, C, device=dev), torch.empty(M, C, device=dev)
    dxm = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    drop = ops.Dropout(0.2, 1, 3)

    def fwd():
        ops.raw_ln_fwd(x, g, b, y, mean, rstd)

    def bwd():
        ops.raw_ln_bwd(dy, x, g, mean, rstd, dres, dx, dg, db, dxm=dxm, dropout=drop)

    fwd()
    return fwd, bwd


LN_BYTES_FWD = M * C * 6.0 + 8.0 * M            # fp32 in, bf16 out, mean / rstd
LN_BYTES_BWD = M * C * (2 + 4 + 4 + 4 + 2.0)    # dy bf16, x, dres in; dx fp32, dxm bf16 out


def adamw_set(dev="cuda", n=10_800_464):
    from drakegpt_b200.optim import FusedAdamW

    class Flat:
        pass
    f = Flat()
    f.device = torch.device(dev)
    f.p, f.g = torch.randn(n, device=dev), torch.randn(n, device=dev)
    f.m, f.v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    f.shadow = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    f.n_live = n
    opt = FusedAdamW(f, lr=3e-4)
    return lambda: opt.launch(zero_grad=True)


ADAMW_BYTES = 10_800_464 * (4 * 4 + 4 * 4 + 2.0)  # p, g, m, v in; p, m, v, g(zero) out; bf16 shadow out


def time_launches(fns, reps=4, replays=5):
    """Mean device microseconds per launch over the functions in `fns` (one launch each), graph-replayed."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns:
                f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (replays * reps * len(fns))


def kernel_family_table(pk):
    """[{kernel, us, bound, achieved, frac}] for every kernel family of the step; pk = bench.peaks()."""
    rows = []
    for name in GEMM_SHAPES:
        us = time_launches([make_gemm(name) for _ in range(R)])
        tf = gemm_flops(name) / us / 1e6
        rows.append({"kernel": "gemm_" + name, "us": round(us, 2), "bound": "tensor", "achieved": round(tf, 1),
                     "unit": "TFLOP/s", "frac": round(tf / pk["bf16_burst"], 3)})
        torch.cuda.empty_cache()
    for name in ("proj_fwd", "ffn2_fwd"):  # the same GEMMs with the next LayerNorm folded in (what the step launches)
        us = time_launches([make_gemm_ln(name) for _ in range(R)])
        gb = gemm_ln_bytes(name) / us / 1e3
        rows.append({"kernel": "gemm_ln_" + name, "us": round(us, 2), "bound": "hbm", "achieved": round(gb, 1), "unit": "GB/s",
                     "frac": round(gb / pk["hbm"], 3), "tflops": round(gemm_flops(name) / us / 1e6, 1)})
        torch.cuda.empty_cache()
    sets = [attn_set(0.2) for _ in range(R)]
    for nm, i, fl in (("attn_fwd", 0, ATTN_FLOPS_FWD), ("attn_bwd", 1, ATTN_FLOPS_BWD)):
        us = time_launches([s[i] for s in sets])
        tf = fl / us / 1e6
        rows.append({"kernel": nm, "us": round(us, 2), "bound": "tensor", "achieved": round(tf, 1), "unit": "TFLOP/s",
                     "frac": round(tf / pk["bf16_burst"], 3),
                     "hbm_frac": round((ATTN_BYTES_FWD if i == 0 else ATTN_BYTES_BWD) / us / 1e3 / pk["hbm"], 3)})
    del sets
    sets = [ln_set() for _ in range(R)]
    for nm, i, by in (("ln_fwd", 0, LN_BYTES_FWD), ("ln_bwd", 1, LN_BYTES_BWD)):
        us = time_launches([s[i] for s in sets])
        gb = by / us / 1e3
        rows.append({"kernel": nm, "us": round(us, 2), "bound": "hbm", "achieved": round(gb, 1), "unit": "GB/s",
                     "frac": round(gb / pk["hbm"], 3)})
    del sets
    us = time_launches([adamw_set() for _ in range(2)], reps=6)
    gb = ADAMW_BYTES / us / 1e3
    rows.append({"kernel": "adamw", "us": round(us, 2), "bound": "hbm", "achieved": round(gb, 1), "unit": "GB/s",
                 "frac": round(gb / pk["hbm"], 3)})
    torch.cuda.empty_cache()
    return rows
