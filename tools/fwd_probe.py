"""Device time of chosen GEMM families exactly as the training step launches them: python tools/fwd_probe.py [names]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import probes  # noqa: E402

for name in (sys.argv[1:] or ["qkv_fwd", "ffn1_fwd", "ffn2_dgrad"]):
    us = probes.time_launches([probes.make_gemm(name) for _ in range(probes.R)])
    print(f"{name}: {us:.2f} us  {probes.gemm_flops(name) / us / 1e6:.0f} TFLOP/s", flush=True)
