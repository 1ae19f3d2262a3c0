#!/bin/bash
# fused residual GEMM + LayerNorm: parity tests, stand-alone timing, training step with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "layernorm_fused or fused_layernorm" 2>&1 | tail -15
timeout 300 python tools/gemm_ln_probe.py 2>&1 | tail -4
for f in 1 0; do echo "== DGPT_FUSE_LN=$f"; DGPT_FUSE_LN=$f timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-kernel-table 2>gpurun_out/bench_fuse$f.err | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['ms_per_step'],4),d['gpu_launches'])" || tail -5 gpurun_out/bench_fuse$f.err; done
