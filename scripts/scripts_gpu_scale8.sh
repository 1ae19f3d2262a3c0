#!/bin/bash
# 8-GPU weak scaling: fused peer-memory optimizer step (default) vs NCCL all-reduce + replicated AdamW
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((29600+RANDOM%100)) bench.py --gpus $NP --steps 100 --warmup 5 2>gpurun_out/scale.err | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['n_gpus'],round(d['value']),round(d['ms_per_step'],3),d['clocks']['reasons'])" || tail -5 gpurun_out/scale.err; }
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline --no-kernel-table 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['n_gpus'],round(d['value']),round(d['ms_per_step'],3))"
for NP in 2 4 8; do [ $NP -gt $N ] && break; run DGPT_DP_MODE=peer; done
NP=$N; run DGPT_DP_MODE=nccl
NP=$N; run DGPT_DP_MODE=peer DGPT_DP_BCAST=all
timeout 600 python -m pytest tests/test_gpu_dp.py -q --timeout 500 -x 2>&1 | tail -3
