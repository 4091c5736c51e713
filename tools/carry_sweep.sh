#!/bin/bash
# (1) step cap of the carried BULK launches (IKB_CARRY_CAP), device-resident arm; (2) host batches carrying too (IKB_QUEUE_CARRY_HOST=1)
run() {  # label, then bench.py arguments
  label=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --no-extras "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('$label steps=%d: value %.1f M solves/s %.4f ms/step frac %.4f | e2e %.1f M (full_io %.1f M)' % (d['steps'], d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], e['value']/1e6, e.get('full_io',{}).get('value',0)/1e6), flush=True)"
}
for st in 20 48; do
  for cap in 1 2 3 4 6 8 16; do IKB_CARRY_CAP=$cap run "carry_cap=$cap" --steps $st --warmup 5; done
done
for st in 20 48; do
  for cfg in "8 4" "12 4" "16 4" "16 8" "24 8" "12 6" "12 3"; do
    set -- $cfg
    IKB_QUEUE_CARRY_HOST=1 run "host-carry depth=$1 merge=$2" --steps $st --warmup 5 --e2e-depth $1 --e2e-merge $2
  done
done
