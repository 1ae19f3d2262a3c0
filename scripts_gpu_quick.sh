#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 2>&1 | tail -8
timeout -k 10 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_models.py -q --timeout 500 2>&1 | tail -8
timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
