#!/bin/bash
timeout -k 10 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_models.py -q --timeout 600 -k "generate or greedy" 2>&1 | tail -3
timeout 900 python tools/decode_bench.py 2>&1 | tail -6
