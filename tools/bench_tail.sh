#!/bin/bash
# A/B of the TAIL (latency) launch of the Cassie kernel: IKB_CASSIE_TAIL=0 thread-per-problem W3, t = team-per-problem.
for dt in f64 f32; do for t in ${TAILS:-0 t}; do
  IKB_CASSIE_TAIL=$t timeout 300 python bench.py --no-cpu-baseline --dtype $dt --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$dt tail=$t  %.1f M solves/s  %.4f ms/step  e2e %.1f M  frac %.4f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac']))"
done; done
