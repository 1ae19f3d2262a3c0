#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -q --timeout 500 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
