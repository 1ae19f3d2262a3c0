#!/bin/bash
# round profile run: bench JSON, per-launch device times of one step (ncu, cold-cache/serialised: compare shares),
# graph-replayed kernel probes, and one `ncu --set full` capture of the roofline kernel (FFN1 GEMM)
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
python tools/kernel_probe.py > gpurun_out/kernel_probe.log 2>&1
python tools/gemm_probe.py --single ffn1_fwd > gpurun_out/gemm_single.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_gemm_r1 python tools/gemm_probe.py --single ffn1_fwd > gpurun_out/ncu_gemm.log 2>&1
cat gpurun_out/bench_r1.json | cut -c1-600; cat gpurun_out/gemm_probe.log gpurun_out/kernel_probe.log
