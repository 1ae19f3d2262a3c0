"""Generation sweep (BASELINE.json config #5): TransformerLM_scaled, KV-cached batched decode, 1xB200.

Prompt zeros((b,1)), 255 new tokens (stays inside the exact-KV regime), sampling on the device.
tokens/s = b*255 / wall (CUDA events), after one warm-up call that captures the per-position graphs.
Also times the CPU oracle (reference semantics: full-window recompute per token) at b=1 and b=16."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drakegpt_b200 import model as M

dev = "cuda"
torch.manual_seed(42)
m = M.TransformerLM(80, 384, 256, 6, 6, 0.2).to(dev).eval()
rows = []
for b in [int(x) for x in (sys.argv[1:] or [1, 4, 16, 64, 256, 1024])]:
    idx = torch.zeros((b, 1), dtype=torch.long, device=dev)
    m.generate(idx, 255, seed=1)  # warm-up / graph capture
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = m.generate(idx, 255, seed=2)
    e1.record(); e1.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    assert out.shape == (b, 256)
    # byte model: weights once per step (bf16) + KV read b*t*9216 + KV write b*9216
    bytes_total = sum(21.6e6 + b * t * 9216 + b * 9216 for t in range(255))
    rows.append({"batch": b, "tokens_per_s": b * 255 / dt, "ms_per_token_step": dt / 255 * 1e3, "model_GBps": bytes_total / dt / 1e9})
    print(json.dumps(rows[-1]), flush=True)
if "--cpu" in os.environ.get("DECODE_BENCH", ""):
    from oracle import drake_oracle as O
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for b, n in ((1, 32), (16, 16)):
        t0 = time.perf_counter()
        O.generate("TransformerLM", sd, torch.zeros((b, 1), dtype=torch.long), n)
        dt = time.perf_counter() - t0
        print(json.dumps({"cpu_oracle_batch": b, "tokens_per_s": b * n / dt, "threads": torch.get_num_threads()}), flush=True)
