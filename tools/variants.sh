#!/bin/bash
# run the C3 bench line with every library under variants/ (GPT_LIB_PATH bypasses the source-hash check)
for f in variants/*.so; do
  GPT_LIB_PATH=$PWD/$f python bench.py --steps 20 --batch-walkers 0 --cpu-evals 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$f', round(d['value']), 'evals/s', round(d['ms_per_step'],4), 'ms/step chunk_ms', round(r['kernel_ms'],4), {k:round(v,1) for k,v in r.get('kernel_us_per_launch',{}).items()})"
done
