#!/bin/bash
# round-2 profile run: per-launch device times of one step (ncu launch list) and one `ncu --set full` capture per kernel family
mkdir -p gpurun_out
cd "$(dirname "$0")"
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table > gpurun_out/bench_nograph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 360 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table > gpurun_out/ncu_launch.log 2>&1
for spec in "gemm_ffn1_dgrad:gemm_tc" "gemm_ffn1_fwd:gemm_tc" "attn_fwd:attn_fwd_tc2" "attn_bwd:attn_bwd_tc" "ln_bwd:ln_bwd_stream" "ln_fwd:ln_fwd_rows" "adamw:adamw_kernel" "lmhead_ce:lmhead_ce" "decode_attn:decode_attn" "decode_persistent:decode_persistent"; do
  t=${spec%%:*}; k=${spec##*:}
  python tools/ncu_target.py $t > gpurun_out/ncu_plain_$t.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_$t python tools/ncu_target.py $t > gpurun_out/ncu_$t.log 2>&1
  echo "$t rc=$?"
done
ls -la gpurun_out/*.ncu-rep
