#!/bin/bash
# 2-GPU data-parallel checks: the fused peer-memory optimizer step vs NCCL, parity and speed
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp.py -q --timeout 500 -x > gpurun_out/dp2_pytest.log 2>&1; tail -3 gpurun_out/dp2_pytest.log
for mode in peer nccl; do
  echo "== DGPT_DP_MODE=$mode"
  DGPT_DP_MODE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%50)) bench.py --gpus 2 --steps 100 --warmup 5 2>gpurun_out/dp2_$mode.err | tail -1 > gpurun_out/dp2_$mode.json
  python -c "import json;d=json.load(open('gpurun_out/dp2_$mode.json'));print(d['n_gpus'],round(d['value']),round(d['ms_per_step'],3),d['e2e']['value'],d['clocks'])" || tail -20 gpurun_out/dp2_$mode.err
done
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline --no-kernel-table 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['n_gpus'],round(d['value']),round(d['ms_per_step'],3))"
