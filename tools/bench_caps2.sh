#!/bin/bash
for t in 0 t; do for cap in ${CAPS:-16 32 48 64}; do
  IKB_CASSIE_TAIL=$t IKB_BULK_CAP=$cap timeout 300 python bench.py --no-cpu-baseline --dtype f64 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('f64 tail=$t cap=$cap  %.1f M solves/s  %.4f ms/step' % (d['value']/1e6, d['ms_per_step']))"
done; done
