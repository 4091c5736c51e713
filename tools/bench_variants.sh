#!/bin/bash
# Runs bench.py for every (dtype, role count) variant of the Cassie kernel; prints value / ms per step.
for dt in f64 f32; do for r in 1 2 3; do
  IKB_CASSIE_ROLES=$r timeout 300 python bench.py --no-cpu-baseline --dtype $dt --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$dt roles=$r  %.1f M solves/s  %.4f ms/step  e2e %.1f M  frac %.4f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac']))"
done; done
