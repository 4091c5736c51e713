#!/bin/bash
# BULK step cap vs queue merge factor (device-resident value only)
for m in ${MERGES:-4 8}; do for cap in ${CAPS:-16 24 32 48}; do
  IKB_BULK_CAP=$cap timeout 300 python bench.py --no-cpu-baseline --merge $m --depth $((2*m)) --steps 24 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('merge=$m cap=$cap  %.1f M solves/s  %.4f ms/step  isolated %.4f  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['config']['isolated_ms_per_batch'], d['e2e']['value']/1e6))"
done; done
