"""Debug probe: per-phase cycle stamps of the GEMM epilogue (needs csrc/libdrakegpt_b200_ts.so, built with -DDGPT_GEMM_TS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drakegpt_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libdrakegpt_b200.so", "libdrakegpt_b200_ts.so")
import torch
from drakegpt_b200 import ops
M, N, K = 16384, 1536, 384
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); bias = torch.zeros(N, device="cuda")
for i in range(2):
    ops.raw_gemm(a, w, out, bias=bias, relu=True)
torch.cuda.synchronize()
