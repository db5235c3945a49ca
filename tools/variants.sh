#!/bin/bash
# run bench lines with every library under variants/ (GPT_LIB_PATH bypasses the source-hash check)
# usage: tools/variants.sh [workload ...]   (default: c3)
wl=${@:-c3}
for f in variants/*.so; do
  for w in $wl; do
  GPT_LIB_PATH=$PWD/$f python bench.py --steps 20 --batch-walkers 0 --cpu-evals 16 --workload $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$f', '$w', round(d['value']), 'evals/s', round(d['ms_per_step'],4), 'ms/step chunk_ms', round(r['kernel_ms'],4), {k:round(v,1) for k,v in r.get('kernel_us_per_launch',{}).items()})"
  done
done
