"""cuBLAS (torch.matmul, bf16) on the model's twelve GEMM shapes, same protocol as tools/gemm_probe.py (6 rotating
operand sets, CUDA-graph replay, CUDA events): the practical ceiling a library GEMM reaches on these small-K shapes.
Measurement aid only -- nothing in drakegpt_b200/ calls cuBLAS."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

M = 16384
shapes = {  # name: (rows, cols, k, form)  form "nt": A[m,k] @ B[n,k]^T ; "nn": A[m,k] @ B[k,n] ; "tn": A[k,m]^T @ B[k,n]
    "qkv_fwd": (M, 1152, 384, "nt"), "proj_fwd": (M, 384, 384, "nt"), "ffn1_fwd": (M, 1536, 384, "nt"),
    "ffn2_fwd": (M, 384, 1536, "nt"), "ffn2_dgrad": (M, 1536, 384, "nn"), "ffn1_dgrad": (M, 384, 1536, "nn"),
    "qkv_dgrad": (M, 384, 1152, "nn"), "proj_dgrad": (M, 384, 384, "nn"), "ffn1_wgrad": (1536, 384, M, "tn"),
    "ffn2_wgrad": (384, 1536, M, "tn"), "qkv_wgrad": (1152, 384, M, "tn"), "proj_wgrad": (384, 384, M, "tn"),
}
R = 6
for name, (m, n, k, form) in shapes.items():
    sets = []
    for _ in range(R):
        if form == "nt":
            a, b = torch.randn(m, k, device="cuda").bfloat16(), torch.randn(n, k, device="cuda").bfloat16()
            f = (lambda a=a, b=b, o=torch.empty(m, n, device="cuda", dtype=torch.bfloat16): torch.matmul(a, b.t(), out=o))
        elif form == "nn":
            a, b = torch.randn(m, k, device="cuda").bfloat16(), torch.randn(k, n, device="cuda").bfloat16()
            f = (lambda a=a, b=b, o=torch.empty(m, n, device="cuda", dtype=torch.bfloat16): torch.matmul(a, b, out=o))
        else:
            a, b = torch.randn(k, m, device="cuda").bfloat16(), torch.randn(k, n, device="cuda").bfloat16()
            f = (lambda a=a, b=b, o=torch.empty(m, n, device="cuda", dtype=torch.bfloat16): torch.matmul(a.t(), b, out=o))
        sets.append(f)
    for f in sets:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            for f in sets:
                f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * 4 * R)
    print(f"cublas {name}: {us:.1f} us  {2.0 * m * n * k / us / 1e6:.0f} TFLOP/s", flush=True)
