#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; s=$(date +%s); timeout -k 10 $TMO "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$? $(( $(date +%s) - s ))s" >> gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
TMO=400 run k_round2 python -m pytest tests/test_gpu_round2.py -q --timeout 300
TMO=400 run kernels python -m pytest tests/test_gpu_kernels.py -q --timeout 300
TMO=600 run models python -m pytest tests/test_gpu_models.py -q --timeout 500
TMO=600 run engine python -m pytest tests/test_gpu_engine.py -q --timeout 500
TMO=600 run cli python -m pytest tests/test_gpu_cli.py -q --timeout 500
TMO=300 run smoke python __graft_entry__.py smoke
TAILN=1 TMO=600 run decode python bench.py --workload decode --warmup 1 --decode-batches 1,4,8,16,32,64,72,128,256,1024
python - <<'PY'
import json
d=json.loads(open('gpurun_out/decode.log').read().strip().splitlines()[-1])
for r in d['sweep']: print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})
print(d['cpu_baseline'])
PY
cat gpurun_out/summary.txt
