#!/bin/bash
# A/B of GEMM scheduling variants at the model's twelve GEMM shapes (device time per launch, graph replay)
mkdir -p gpurun_out
out=gpurun_out/gemm_ab.log; : > $out
for v in "DGPT_GEMM_ROLE_HI=0" "DGPT_GEMM_ROLE_HI=1" "DGPT_GEMM_ROLE_HI=0 DGPT_GEMM_CTA_GROUP=2" "DGPT_GEMM_ROLE_HI=1 DGPT_GEMM_CTA_GROUP=2" \
         "DGPT_GEMM_ROLE_HI=0 DGPT_GEMM_DEBUG=1" "DGPT_GEMM_ROLE_HI=1 DGPT_GEMM_DEBUG=1" "DGPT_GEMM_ROLE_HI=1 DGPT_GEMM_CTA_GROUP=2 DGPT_GEMM_DEBUG=1" \
         "DGPT_GEMM_ROLE_HI=1 DGPT_GEMM_CTA_GROUP=2 DGPT_GEMM_DEBUG=4"; do
  echo "== $v" >> $out
  env $v timeout 200 python tools/gemm_probe.py >> $out 2>&1
done
cat $out
