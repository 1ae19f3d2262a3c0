#!/bin/bash
# decode attention: parity test + stand-alone timing, default policy vs pinned warps per (sequence, head)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -q -x -k "decode or generation" 2>&1 | tail -3
for w in ${WPUS:-0}; do echo "== DGPT_DECODE_WPU=$w"; DGPT_DECODE_WPU=$w timeout 200 python tools/decode_attn_bench.py 1024 256 64 16; done 2>&1 | tee gpurun_out/decattn.log
timeout 300 python bench.py --workload decode --decode-batches 16,64,256,1024 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/decode_sweep.json; python -c "
import json
d=json.load(open('gpurun_out/decode_sweep.json'))
for r in d['sweep']: print({k: round(v, 3) for k, v in r.items()})"
timeout 200 python tools/decode_profile.py 2>&1 | tail -12
