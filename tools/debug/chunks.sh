#!/bin/bash
for mb in $1; do
  AKSHAR_CHUNK_MB=$mb timeout 300 python bench.py --mb 1024 --steps 3 --warmup 3 --cpu-sample-mb 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $mb MB: value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'GB/s', round(d['e2e']['ms_per_step'],1),'ms')"
done
