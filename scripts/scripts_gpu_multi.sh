#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_n$n.json
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550+n)) bench.py --gpus $n --steps 200 --warmup 5 2>gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_n$n.json; fi
  python -c "import json;d=json.load(open('gpurun_out/scale_n$n.json'));print(d['n_gpus'],round(d['value']),round(d['ms_per_step'],3),d['clocks'])" || tail -5 gpurun_out/scale_n$n.err
done
