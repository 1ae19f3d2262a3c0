#!/bin/bash
timeout -k 10 900 python -m pytest tests/test_gpu_models.py -q --timeout 600 -k "train_curve" 2>&1 | grep -E "assert|Error|passed|failed|tensor" | head -20
