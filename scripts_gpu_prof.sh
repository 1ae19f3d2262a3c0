#!/bin/bash
mkdir -p gpurun_out
python tools/ln_probe.py > gpurun_out/ln_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ln_fwd|ln_bwd_fast|colsum" -s 9 -c 3 -o gpurun_out/prof_ln python tools/ln_probe.py > gpurun_out/ncu_ln.log 2>&1
tail -3 gpurun_out/ncu_ln.log
