#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; s=$(date +%s); timeout -k 10 $TMO "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$? $(( $(date +%s) - s ))s" >> gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
TMO=300 run gen_paths python -m pytest tests/test_gpu_round2.py -q --timeout 250 -x -k "generation_paths or decode_attention"
TMO=300 run engine python -m pytest tests/test_gpu_engine.py -q --timeout 250 -k "generate"
out=gpurun_out/exp7.log; : > $out
for c in 16 8; do echo "== persistent decode, cluster of $c" >> $out; DGPT_DECODE_CLUSTER=$c timeout 200 python bench.py --workload decode --no-cpu-baseline --warmup 1 --decode-batches 1,8,64 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['sweep']: print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})" >> $out 2>&1; done
DGPT_CLOCK_PROBE=1 timeout 200 python tools/clock_probe.py decode >> $out 2>&1
cat gpurun_out/summary.txt $out
