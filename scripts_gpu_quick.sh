#!/bin/bash
timeout -k 10 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1.json; cut -c1-250 gpurun_out/bench_r1.json
