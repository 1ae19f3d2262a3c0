#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; s=$(date +%s); timeout -k 10 $TMO "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$? $(( $(date +%s) - s ))s" >> gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
TMO=200 run k_round2 python -m pytest tests/test_gpu_round2.py -q --timeout 120 -x
TMO=200 run cli python -m pytest tests/test_gpu_cli.py -q --timeout 180
out=gpurun_out/exp3.log; : > $out
DGPT_CLOCK_PROBE=1 timeout 300 python tools/clock_probe.py >> $out 2>&1
for v in "DGPT_GEMM_FORCE_BN=128" "DGPT_GEMM_FORCE_BN=128 DGPT_GEMM_DEBUG=1" "DGPT_GEMM_BN192=0 DGPT_GEMM_FORCE_BN=128"; do
  echo "== $v" >> $out
  env $v timeout 200 python tools/gemm_probe.py >> $out 2>&1
done
cat gpurun_out/summary.txt $out
