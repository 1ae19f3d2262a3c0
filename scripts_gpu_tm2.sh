#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -q -x -k "256_row" --timeout 200 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "tcgen05" --timeout 200 2>&1 | tail -3
echo "== TM=1"; DGPT_GEMM_TM=1 timeout 200 python tools/gemm_probe.py qkv_fwd ffn1_fwd ffn2_dgrad ffn1_dgrad qkv_dgrad proj_dgrad
echo "== TM=2"; timeout 200 python tools/gemm_probe.py qkv_fwd ffn1_fwd ffn2_dgrad ffn1_dgrad qkv_dgrad proj_dgrad
echo "== TM=2 no epilogue"; DGPT_GEMM_DEBUG=1 timeout 200 python tools/gemm_probe.py qkv_fwd ffn1_fwd ffn2_dgrad ffn1_dgrad qkv_dgrad proj_dgrad
DGPT_CLOCK_PROBE=1 timeout 200 python tools/clock_probe.py qkv_fwd ffn1_fwd ffn2_dgrad ffn1_dgrad qkv_dgrad proj_dgrad
