#!/bin/bash
# Small-batch latency (one launch in the latency configuration): thread-per-problem W3 (tail=0) vs team kernel (tail=t).
for dt in f64 f32; do for B in ${BATCHES:-1024 4096 9472}; do for t in 0 t; do
  IKB_CASSIE_TAIL=$t timeout 300 python bench.py --no-cpu-baseline --dtype $dt --batch $B --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$dt B=$B tail=$t  %.2f M solves/s  %.4f ms/step' % (d['value']/1e6, d['ms_per_step']))"
done; done; done
