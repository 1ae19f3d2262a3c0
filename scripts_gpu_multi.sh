#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout -k 10 600 python -m pytest tests/test_gpu_dp.py -q --timeout 600 2>&1 | tail -15
for n in 1 2; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_n$n.err | tail -1 > gpurun_out/bench_n$n.json
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 100 --warmup 5 2>gpurun_out/bench_n$n.err | tail -1 > gpurun_out/bench_n$n.json; fi
  cut -c1-220 gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
done
