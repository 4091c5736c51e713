#!/bin/bash
# Sweeps the BULK step cap (IKB_BULK_CAP) for the default kernel choice.
for dt in f64 f32; do for cap in 2 4 6 8 12 16; do
  IKB_BULK_CAP=$cap timeout 300 python bench.py --no-cpu-baseline --dtype $dt --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$dt cap=$cap  %.1f M solves/s  %.4f ms/step  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"
done; done
