#!/bin/bash
# device-resident arm of bench.py against the queue's merge / depth and the BULK step cap, at the driver's step count
for cfg in "8 24 16" "8 24 8" "8 24 4" "8 24 24" "7 21 16" "5 20 16" "5 20 8" "4 16 16" "4 16 8" "4 24 16" "6 24 16" "8 16 16" "8 32 16"; do
  set -- $cfg
  IKB_BULK_CAP=$3 timeout 300 python bench.py --no-cpu-baseline --no-extras --steps ${STEPS:-20} --warmup 5 --merge $1 --depth $2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('steps=%d merge=$1 depth=$2 cap=$3: value %.1f M solves/s  %.4f ms/step  frac %.4f  lone %.3f ms  e2e %.1f M' % (d['steps'], d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['details']['isolated_ms_per_batch'], d['e2e']['value']/1e6), flush=True)"
done
for cfg in "8 4" "8 2" "6 3" "6 2" "10 5" "4 2" "8 1"; do
  set -- $cfg
  timeout 300 python bench.py --no-cpu-baseline --no-extras --steps ${STEPS:-20} --warmup 5 --e2e-depth $1 --e2e-merge $2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('e2e depth=$1 merge=$2 steps=%d: e2e %.1f M solves/s (full_io %.1f M)' % (e['steps'], e['value']/1e6, e.get('full_io',{}).get('value',0)/1e6), flush=True)"
done
