#!/bin/bash
python tools/ln_probe.py 2>&1 | tail -4
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q --timeout 500 -x 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
