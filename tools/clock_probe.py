"""Effective SM clock and per-phase cycle stamps of the GEMM / attention kernels under steady load.

    DGPT_CLOCK_PROBE=1 python tools/clock_probe.py

Each case is replayed from a CUDA graph (6 rotating buffer sets) long enough for the power controller to settle; the
kernels stamp clock64 and %globaltimer in CTA 0 (common.cuh: clock_probe_*), so cycles / ns of the LAST launch give
the SM clock the kernel really ran at -- nvidia-smi's 100 ms samples cannot see it."""
import ctypes as C
import os
import sys

os.environ.setdefault("DGPT_CLOCK_PROBE", "1")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import probes  # noqa: E402
from drakegpt_b200 import _lib  # noqa: E402


def stamps():
    buf = (C.c_uint64 * 64)()
    _lib.check(_lib.lib().dgpt_debug_clock_stamps(buf), "clock stamps")
    return list(buf)


def report(name, us):
    s = stamps()
    cyc, ns = s[2] - s[0], s[3] - s[1]
    print(f"{name}: {us:.1f} us/launch | CTA0 {cyc} cycles in {ns} ns -> {cyc / max(ns, 1) * 1e3:.0f} MHz", flush=True)
    return s


want = sys.argv[1:]
for name in probes.GEMM_SHAPES:
    if want and not any(w in name for w in want):
        continue
    us = probes.time_launches([probes.make_gemm(name) for _ in range(probes.R)], reps=8, replays=10)
    s = report("gemm_" + name, us)
    mm, ep, pr = s[8:11], s[16:21], s[24:26]
    if mm[0] and ep[0]:
        print(f"  issuer: {mm[0]} cycles, waits: accumulator free {100 * mm[1] / mm[0]:.0f} %, operands {100 * mm[2] / mm[0]:.0f} % | "
              f"epilogue warp 0: {ep[0]} cycles, waits: accumulator ready {100 * ep[1] / ep[0]:.0f} %, staging tile free "
              f"{100 * ep[2] / ep[0]:.0f} %, residual {100 * ep[3] / ep[0]:.0f} %, final drain {ep[4]} cycles | producer: "
              f"{pr[0]} cycles, ring full {100 * pr[1] / max(pr[0], 1):.0f} %")
    torch.cuda.empty_cache()
if not want or any("attn" in w for w in want):
    sets = [probes.attn_set(0.2) for _ in range(probes.R)]
    us = probes.time_launches([s[0] for s in sets], reps=8, replays=10)
    s = report("attn_fwd", us)
    t0 = s[0]
    for k in range(3):
        w = s[4 + 8 * k: 4 + 8 * k + 6]
        if w[0]:
            print(f"  CTA0 WG0 item {2 * k}: start +{w[0] - t0} | S wait {w[1] - w[0]} | max pass {w[2] - w[1]} | exp pass {w[3] - w[2]}"
                  f" | O wait {w[4] - w[3]} | output {w[5] - w[4]}")
if not want or any("decode" in w for w in want):
    from drakegpt_b200 import model as M
    torch.manual_seed(0)
    m = M.TransformerLM(80, 384, 256, 6, 6, 0.2).to("cuda").eval()
    for b in (1, 8):
        idx = torch.zeros((b, 1), dtype=torch.long, device="cuda")
        m.generate(idx, 255, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.generate(idx, 255, seed=2)
        e1.record()
        e1.synchronize()
        s = stamps()
        ph = [s[i + 1] - s[i] for i in range(4, 63) if s[i + 1] > s[i] > 0]
        names = ["qkv", "attn", "proj", "ffn1", "ffn2"] * 6 + ["lm_head", "sample"]
        print(f"persistent decode b={b}: {e0.elapsed_time(e1) * 1e3 / 255:.1f} us/token; phases of the last position (cycles):")
        print("  " + ", ".join(f"{n} {c}" for n, c in zip(names, ph)))
        q = s[48:53]
        if q[0]:
            print(f"  proj phase of layer 1, thread 0: gemv + emit {q[1] - q[0]} | prefetch issue {q[2] - q[1]} | "
                  f"arrive {q[3] - q[2]} | wait {q[4] - q[3]}")
