"""Times dgpt_decode_attn alone at generation batch sizes (KV cache larger than L2, so every launch streams from HBM).

usage: python tools/decode_attn_bench.py [batch ...]      env DGPT_DECODE_WPU=1|2|4|8 pins the warps per (sequence, head)
Prints, per (batch, keys): us per launch and the KV bytes / time against the measured HBM peak.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drakegpt_b200 import ops  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [1024, 256, 64, 16]
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6532.9
    D, NH, H, ctx = 384, 6, 64, 256
    dev = torch.device("cuda", 0)
    for B in batches:
        cache = (torch.randn(B, ctx, 3 * D, device=dev) * 0.5).bfloat16()
        out = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
        for nk in (16, 32, 64, 128, 192, 255):
            t = nk - 1
            fn = lambda: ops.raw_decode_attn(cache[:, t:t + 1, :D], cache[:, :nk, D:2 * D], cache[:, :nk, 2 * D:], out.view(B, 1, D), NH, H, 0.125)  # noqa: E731
            for _ in range(3):
                fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10):
                    fn()
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            byts = B * nk * 2 * D * 2
            print(f"batch {B:5d} keys {nk:3d}: {us:7.2f} us  {byts / us / 1e3:7.1f} GB/s  ({byts / us / 1e3 / peak:.2f} of HBM peak)", flush=True)


if __name__ == "__main__":
    main()
