#!/bin/bash
# GPU check: suites in separate processes (a hang in one cannot hide the others), then a short bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; s=$(date +%s); timeout -k 10 $TMO "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$? $(( $(date +%s) - s ))s" >> gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TMO=300 run k_round2 python -m pytest tests/test_gpu_round2.py -q --timeout 280 -x -k "not 200_steps"
TMO=400 run curves python -m pytest tests/test_gpu_round2.py -q --timeout 380 -k "200_steps"
TMO=300 run k_attn_tc python -m pytest tests/test_gpu_kernels.py -q -k "attention_tcgen05" --timeout 200
TMO=600 run kernels python -m pytest tests/test_gpu_kernels.py -q -k "not attention_tcgen05" --timeout 300
TMO=900 run models python -m pytest tests/test_gpu_models.py -q --timeout 600
TMO=900 run engine python -m pytest tests/test_gpu_engine.py -q --timeout 600
TMO=900 run cli python -m pytest tests/test_gpu_cli.py -q --timeout 600
TMO=600 run smoke python __graft_entry__.py smoke
TAILN=1 TMO=900 run bench python bench.py --steps 50 --warmup 5
TAILN=1 TMO=300 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
cat gpurun_out/summary.txt
