#!/bin/bash
# 8-GPU A/B of the gradient all-reduce schedule (one message at the end of backward vs overlapped buckets)
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600+RANDOM%100)) bench.py --gpus 8 --steps 100 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['ms_per_step'],3))"; }
run DGPT_DP_OVERLAP=0
run DGPT_DP_OVERLAP=1
run DGPT_DP_OVERLAP=0
