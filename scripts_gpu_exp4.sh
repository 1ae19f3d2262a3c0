#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; s=$(date +%s); timeout -k 10 $TMO "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$? $(( $(date +%s) - s ))s" >> gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
TMO=120 run decode_attn python -m pytest tests/test_gpu_round2.py -q --timeout 100 -x -k "decode_attention"
TMO=300 run gen_paths python -m pytest tests/test_gpu_round2.py -q --timeout 250 -x -k "generation_paths"
TMO=200 run k_round2 python -m pytest tests/test_gpu_round2.py -q --timeout 120 -k "not generation_paths and not decode_attention"
TMO=200 run k_attn python -m pytest tests/test_gpu_kernels.py -q --timeout 120 -k "attention"
TMO=400 run engine python -m pytest tests/test_gpu_engine.py -q --timeout 300
out=gpurun_out/exp4.log; : > $out
DGPT_CLOCK_PROBE=1 timeout 300 python tools/clock_probe.py attn >> $out 2>&1
echo "== staging bufs 2 (default)" >> $out; timeout 200 python tools/gemm_probe.py >> $out 2>&1
echo "== staging bufs 1 (one more ring stage)" >> $out; DGPT_LIB=$PWD/drakegpt_b200/csrc/libdrakegpt_b200_sb1.so timeout 200 python tools/gemm_probe.py >> $out 2>&1
echo "== staging bufs 1, no epilogue" >> $out; DGPT_GEMM_DEBUG=1 DGPT_LIB=$PWD/drakegpt_b200/csrc/libdrakegpt_b200_sb1.so timeout 200 python tools/gemm_probe.py >> $out 2>&1
echo "== decode sweep" >> $out
TAILN=1 TMO=600 run decode python bench.py --workload decode --no-cpu-baseline --warmup 1
echo "== decode sweep, persistent off" >> $out
DGPT_DECODE_PERSISTENT=0 TAILN=1 TMO=300 run decode_np python bench.py --workload decode --no-cpu-baseline --warmup 1 --decode-batches 1,4
TAILN=1 TMO=300 run bench python bench.py --steps 50 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt $out
