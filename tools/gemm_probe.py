"""Run the tcgen05 GEMM at the scaled-model shapes a few times (ncu target + quick timing)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drakegpt_b200 import ops
from drakegpt_b200._lib import MAJOR_K, MAJOR_MN

dev = "cuda"
shapes = {  # name: (M, N, K, a_major, b_major, out dtype, epilogue)
    "ffn1_fwd": (16384, 1536, 384, MAJOR_K, MAJOR_K, torch.bfloat16, "bias_relu"),
    "ffn2_fwd": (16384, 384, 1536, MAJOR_K, MAJOR_K, torch.float32, "bias_drop_res"),
    "qkv_fwd": (16384, 1152, 384, MAJOR_K, MAJOR_K, torch.bfloat16, "none"),
    "ffn2_dgrad": (16384, 1536, 384, MAJOR_K, MAJOR_MN, torch.bfloat16, "relu_aux"),
    "ffn1_wgrad": (1536, 384, 16384, MAJOR_MN, MAJOR_MN, torch.float32, "splitk"),
}
which = sys.argv[1:] or list(shapes)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for name in which:
    M, N, K, am, bm, odt, epi = shapes[name]
    A = torch.randn((M, K) if am == MAJOR_K else (K, M), device=dev).bfloat16()
    B = torch.randn((N, K) if bm == MAJOR_K else (K, N), device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=odt)
    kw = {}
    if epi == "bias_relu":
        kw = dict(bias=torch.zeros(N, device=dev), relu=True)
    elif epi == "bias_drop_res":
        kw = dict(bias=torch.zeros(N, device=dev), residual=torch.zeros(M, N, device=dev), dropout=ops.Dropout(0.2, 1, 1))
    elif epi == "relu_aux":
        kw = dict(relu_aux=torch.randn(M, N, device=dev).bfloat16())
    elif epi == "splitk":
        kw = dict(accumulate=True, split_k=4)
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.raw_gemm(A, B, out, a_major=am, b_major=bm, **kw)
        e1.record(); e1.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    t = statistics.mean(ts)
    print(f"{name}: {t:.1f} us  {2.0*M*N*K/t/1e6:.0f} TFLOP/s", flush=True)
