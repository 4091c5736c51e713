#!/bin/bash
# e2e arm of bench.py against the host queue's depth / merge (and the step count, which sets the weight of fill / drain)
for cfg in "8 4" "12 4" "16 4" "12 6" "16 8" "24 8" "32 8"; do
  set -- $cfg
  for steps in ${STEPS:-48}; do
    timeout 300 python bench.py --no-cpu-baseline --no-extras --steps $steps --e2e-depth $1 --e2e-merge $2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('e2e depth=$1 merge=$2 steps=%d: e2e %.1f M solves/s (full_io %.1f M)   value %.1f M' % (e['steps'], e['value']/1e6, e.get('full_io',{}).get('value',0)/1e6, d['value']/1e6), flush=True)"
  done
done
