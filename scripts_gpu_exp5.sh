#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; s=$(date +%s); timeout -k 10 $TMO "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$? $(( $(date +%s) - s ))s" >> gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
TMO=300 run gen_paths python -m pytest tests/test_gpu_round2.py -q --timeout 250 -x -k "generation_paths or decode_attention"
TMO=300 run engine python -m pytest tests/test_gpu_engine.py -q --timeout 250 -k "generate"
TMO=300 run kernels python -m pytest tests/test_gpu_kernels.py -q --timeout 250 -k "tcgen05"
out=gpurun_out/exp5.log; : > $out
for c in 148 74 48 32; do echo "== persistent decode, $c CTAs" >> $out; DGPT_DECODE_CTAS=$c timeout 200 python bench.py --workload decode --no-cpu-baseline --warmup 1 --decode-batches 1,8 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['sweep']: print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})" >> $out 2>&1; done
echo "== decode step profile b=1024 t=128" >> $out; timeout 200 python tools/decode_profile.py 1024 128 >> $out 2>&1
echo "== decode step profile b=64 t=128" >> $out; timeout 200 python tools/decode_profile.py 64 128 >> $out 2>&1
echo "== gemm, TMA stores" >> $out; timeout 200 python tools/gemm_probe.py >> $out 2>&1
echo "== gemm, coalesced LSU stores" >> $out; DGPT_GEMM_STORE=coalesced timeout 200 python tools/gemm_probe.py >> $out 2>&1
DGPT_GEMM_STORE=coalesced TMO=300 run kernels_coal python -m pytest tests/test_gpu_kernels.py -q --timeout 250 -k "tcgen05"
cat gpurun_out/summary.txt $out
