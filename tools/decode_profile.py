"""Per-kernel device time of ONE launch-per-kernel decode step (batch b, position t) of TransformerLM_scaled:
every ops.raw_* launch of Runner._decode_token + LM head + sampling is bracketed by CUDA events (eager, after a
warm-up pass), grouped by kernel kind.  `python tools/decode_profile.py [b] [t]`"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from drakegpt_b200 import model as M, ops  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
m = M.TransformerLM(80, 384, 256, 6, 6, 0.2).to("cuda").eval()
r = m.runner()
r.flat.refresh_shadow()
caches = [torch.randn(b, 256, 3 * 384, device="cuda").bfloat16() for _ in range(6)]
toks = torch.zeros(b, dtype=torch.long, device="cuda")
names = ["raw_gemm", "raw_ln_fwd", "raw_embed_fwd", "raw_decode_attn", "raw_attn_fwd", "raw_dropout_scale", "raw_sample"]
times = defaultdict(list)
orig = {n: getattr(ops, n) for n in names}


def wrap(n):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig[n](*a, **k)
        e1.record()
        times[n].append((e0, e1))
        return out
    return f


for rep in range(3):
    if rep == 2:
        for n in names:
            setattr(ops, n, wrap(n))
    x = r._decode_token(toks, t, caches)
    logits = r.buf("d.logits", (b, 80), torch.float32)
    xin = x if x.dtype == r.at else ops.raw_dropout_scale(x, r.buf("d.xl", x.shape))
    ops.raw_gemm(xin, r.w("lm_head.weight"), logits, bias=r.f("lm_head.bias"))
    ops.raw_sample(logits, toks, 0, False, 1, t)
torch.cuda.synchronize()
total = 0.0
for n, evs in times.items():
    us = sum(a.elapsed_time(c) for a, c in evs) * 1e3
    total += us
    print(f"{n}: {len(evs)} launches, {us:.1f} us total, {us / len(evs):.1f} us each")
kv = b * (t + 1) * 9216
print(f"batch {b} position {t}: {total:.1f} us in kernels (eager, includes launch gaps); KV read {kv / 1e6:.1f} MB "
      f"-> {kv / 6.5e6:.1f} us at 6.5 TB/s")
