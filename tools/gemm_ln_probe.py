"""Device time of the fused residual GEMM + LayerNorm kernel against the two kernels it replaces (B=64 shapes).
`python tools/gemm_ln_probe.py`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import probes  # noqa: E402

for name in ("proj_fwd", "ffn2_fwd"):
    fused = probes.time_launches([probes.make_gemm_ln(name) for _ in range(probes.R)])
    gemm = probes.time_launches([probes.make_gemm(name) for _ in range(probes.R)])
    ln = probes.time_launches([probes.ln_set()[0] for _ in range(probes.R)])
    print(f"{name}: fused {fused:.2f} us ({probes.gemm_ln_bytes(name) / fused / 1e3:.0f} GB/s algorithmic) | "
          f"gemm {gemm:.2f} + ln_fwd {ln:.2f} = {gemm + ln:.2f} us", flush=True)
