#!/bin/bash
# one `ncu --set full` capture of one kernel family, summarised on the box: bash scripts_gpu_ncu1.sh <target> <kernel regex> [stall lines]
mkdir -p gpurun_out
t=$1; k=$2; n=${3:-40}
python tools/ncu_target.py $t > gpurun_out/ncu_plain_$t.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/r2_$t python tools/ncu_target.py $t > gpurun_out/ncu_$t.log 2>&1
rc=$?
{ echo "# ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 python tools/ncu_target.py $t   (rc=$rc)";
  python tools/ncu_summary.py metrics /tmp/r2_$t.ncu-rep; python tools/ncu_summary.py stalls /tmp/r2_$t.ncu-rep $n; } > gpurun_out/r2_ncu_$t.txt 2>&1
cat gpurun_out/r2_ncu_$t.txt
