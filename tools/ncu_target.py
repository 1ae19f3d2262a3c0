"""One kernel family launched a few times, eagerly, for `ncu -k regex:... -s 2 -c 1` captures.
    python tools/ncu_target.py gemm_ffn1_dgrad | gemm_ffn1_fwd | gemm_ln_proj_fwd | gemm_ln_ffn2_fwd | attn_fwd | attn_bwd | ln_fwd | ln_bwd | adamw |
                               lmhead_ce | decode_attn | decode_persistent
L2 is flushed between launches (a 256 MB memset), so the captured launch reads its operands from HBM."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import probes  # noqa: E402
from drakegpt_b200 import ops  # noqa: E402

what = sys.argv[1]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
if what.startswith("gemm_ln_"):
    fn = probes.make_gemm_ln(what[8:])
elif what.startswith("gemm_"):
    fn = probes.make_gemm(what[5:])
elif what in ("attn_fwd", "attn_bwd"):
    f, b = probes.attn_set(0.2)
    fn = f if what == "attn_fwd" else b
elif what in ("ln_fwd", "ln_bwd"):
    f, b = probes.ln_set()
    fn = f if what == "ln_fwd" else b
elif what == "adamw":
    fn = probes.adamw_set()
elif what == "lmhead_ce":
    x = torch.randn(probes.M, probes.C, device="cuda").bfloat16()
    w = torch.randn(80, probes.C, device="cuda").bfloat16()
    bias, tgt = torch.zeros(80, device="cuda"), torch.randint(0, 80, (probes.M,), device="cuda")
    loss, dl = torch.zeros(1, device="cuda"), torch.zeros(probes.M, 80, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.raw_lmhead_ce(x, w, bias, tgt, loss, dl, None, None)  # noqa: E731
elif what == "decode_attn":
    b, t, D = 1024, 128, 384
    cache = torch.randn(b, 256, 3 * D, device="cuda").bfloat16()
    out = torch.empty(b, 1, D, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.raw_decode_attn(cache[:, t:t + 1, :D], cache[:, :t + 1, D:2 * D], cache[:, :t + 1, 2 * D:], out, 6, 64, 0.125)  # noqa: E731
elif what == "decode_persistent":
    from drakegpt_b200 import model as M
    torch.manual_seed(0)
    m = M.TransformerLM(80, 384, 256, 6, 6, 0.2).to("cuda").eval()
    idx = torch.zeros((1, 1), dtype=torch.long, device="cuda")
    fn = lambda: m.generate(idx, 63, seed=1)  # noqa: E731
else:
    raise SystemExit(f"unknown target {what}")
for _ in range(4):
    flush.zero_()
    fn()
torch.cuda.synchronize()
print("ok", what)
