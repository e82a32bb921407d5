#!/bin/bash
# usage: tools/debug/variants.sh "var_2_4 var_3_4 ..." [bench args]
for v in $1; do
  AKSHAR_B200_LIB=$PWD/akshar_b200/lib/$v.so timeout 300 python bench.py --mb 256 --steps 3 --warmup 3 --cpu-sample-mb 1 ${@:2} 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],2), 'GB/s', d['roofline']['kernels_ms'])"
done
