"""Device time of the four wgrad GEMMs of a block exactly as the training step launches them (tools/probes.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import probes  # noqa: E402

for name in ("ffn1_wgrad", "ffn2_wgrad", "qkv_wgrad", "proj_wgrad"):
    us = probes.time_launches([probes.make_gemm(name) for _ in range(probes.R)])
    print(f"{name}: {us:.2f} us  {probes.gemm_flops(name) / us / 1e6:.0f} TFLOP/s", flush=True)
