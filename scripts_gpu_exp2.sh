#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; s=$(date +%s); timeout -k 10 $TMO "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$? $(( $(date +%s) - s ))s" >> gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
TMO=300 run k_round2 python -m pytest tests/test_gpu_round2.py -q --timeout 280
TMO=300 run k_attn_tc python -m pytest tests/test_gpu_kernels.py -q -k "attention" --timeout 200
TMO=600 run engine python -m pytest tests/test_gpu_engine.py -q --timeout 600
TMO=600 run cli python -m pytest tests/test_gpu_cli.py -q --timeout 600 -k "evaluate_loss"
out=gpurun_out/exp2.log; : > $out
for v in "DGPT_ATTN_FWD=1" "DGPT_ATTN_FWD=2"; do echo "== $v" >> $out; env $v timeout 200 python tools/kernel_probe.py attn_fwd >> $out 2>&1; done
for v in "DGPT_GEMM_DEBUG=1" "DGPT_GEMM_DEBUG=17" "DGPT_GEMM_DEBUG=33" "DGPT_GEMM_DEBUG=16" "DGPT_GEMM_DEBUG=32"; do
  echo "== $v" >> $out
  env $v timeout 200 python tools/gemm_probe.py qkv_fwd ffn1_fwd ffn2_fwd ffn1_dgrad ffn1_wgrad >> $out 2>&1
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv -lms 100 > gpurun_out/clk_probe.csv &
SMI=$!
python tools/gemm_probe.py ffn1_fwd ffn1_fwd ffn1_fwd ffn1_fwd >> $out 2>&1
kill $SMI
sort gpurun_out/clk_probe.csv | uniq -c | sort -rn | head -8 >> $out
cat gpurun_out/summary.txt $out
