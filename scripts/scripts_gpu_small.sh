#!/bin/bash
# configs #1-#3 (small models) + reference arms
for w in bigram singlehead residual; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 10 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['metric'],round(d['value']),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'cpu',round(d['cpu_baseline']['value']),d['cpu_baseline']['kind'])"
  timeout 300 python bench.py --impl reference --workload $w --steps 100 --warmup 5 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('  reference arm:',d['metric'],round(d['value']),d['cpu_baseline']['kind'],d['cpu_baseline']['cores'])"
done
timeout 300 python bench.py --impl reference --workload decode 2>&1 | tail -1 | cut -c1-400
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
