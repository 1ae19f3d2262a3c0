#!/bin/bash
timeout -k 10 200 python -m pytest tests/test_gpu_kernels.py -q -k "tcgen05" --timeout 120 -x 2>&1 | tail -6
echo "--- cta_group=2"; timeout 120 python tools/gemm_probe.py 2>&1 | tail -5
echo "--- cta_group=1"; DGPT_GEMM_CTA_GROUP=1 timeout 120 python tools/gemm_probe.py 2>&1 | tail -5
for d in 1 2; do echo "--- cg=2 debug=$d"; DGPT_GEMM_DEBUG=$d timeout 120 python tools/gemm_probe.py ffn1_fwd ffn2_fwd 2>&1 | tail -2; done
