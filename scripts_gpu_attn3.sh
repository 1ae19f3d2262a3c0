#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -q -x -k "attention_tcgen05" --timeout 200 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" --timeout 200 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_engine.py -q -x --timeout 250 2>&1 | tail -3
for v in 1 2 3; do echo "== DGPT_ATTN_FWD=$v"; DGPT_ATTN_FWD=$v timeout 200 python tools/kernel_probe.py attn_fwd; done
DGPT_CLOCK_PROBE=1 timeout 200 python tools/clock_probe.py attn
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-kernel-table | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['step_tensor_frac']['of_burst'])"
