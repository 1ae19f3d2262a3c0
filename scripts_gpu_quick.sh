#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -q -k "tcgen05" --timeout 200 2>&1 | tail -4
echo "--- normal"; python tools/gemm_probe.py 2>&1 | tail -5
for d in 1 2; do echo "--- debug=$d"; DGPT_GEMM_DEBUG=$d python tools/gemm_probe.py ffn1_fwd qkv_fwd 2>&1 | tail -2; done
