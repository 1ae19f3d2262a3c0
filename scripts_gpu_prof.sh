#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_gemm_r1 python tools/gemm_probe.py ffn1_fwd > gpurun_out/ncu_gemm.log 2>&1
cat gpurun_out/gemm_probe.log; python tools/attn_probe.py; python tools/ln_probe.py
