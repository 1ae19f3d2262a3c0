#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_probe.py ffn1_fwd > gpurun_out/gemm_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_gemm2 python tools/gemm_probe.py ffn1_fwd > gpurun_out/ncu_gemm.log 2>&1
cat gpurun_out/gemm_probe.log
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/bench_nograph.log | cut -c1-200
