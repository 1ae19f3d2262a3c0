#!/bin/bash
# round-2 profile run: per-launch device times of one step (ncu launch list) and one `ncu --set full` capture per kernel
# family, summarised ON the GPU box (tools/ncu_summary.py) so that only text comes back (gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table > gpurun_out/bench_nograph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 360 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table > gpurun_out/ncu_launch.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1
ALL="gemm_ffn1_dgrad:gemm_tc gemm_ffn1_fwd:gemm_tc gemm_ln_proj_fwd:gemm_res_ln gemm_ln_ffn2_fwd:gemm_res_ln attn_fwd:attn_fwd_tc2 attn_bwd:attn_bwd_tc ln_bwd:ln_bwd_stream ln_fwd:ln_fwd_rows adamw:adamw_kernel lmhead_ce:lmhead_ce decode_attn:decode_attn decode_persistent:decode_persistent"
for spec in ${TARGETS:-$ALL}; do   # TARGETS="name:kernel-regex ..." re-captures a subset
  t=${spec%%:*}; k=${spec##*:}
  python tools/ncu_target.py $t > gpurun_out/ncu_plain_$t.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/r2_$t python tools/ncu_target.py $t > gpurun_out/ncu_$t.log 2>&1
  rc=$?
  { echo "# ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 python tools/ncu_target.py $t   (rc=$rc)";
    python tools/ncu_summary.py metrics /tmp/r2_$t.ncu-rep; python tools/ncu_summary.py stalls /tmp/r2_$t.ncu-rep 18; } > gpurun_out/r2_ncu_$t.txt 2>&1
  echo "$t rc=$rc"
done
cuobjdump -sass drakegpt_b200/csrc/libdrakegpt_b200.so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | grep -E "UTC|UTMA|UBLKCP|LDTM|STTM|HMMA|SYNCS|UCGABAR|MUFU|REDG|ATOM" > gpurun_out/r2_sass_histogram.txt
ls -la gpurun_out | tail -30
