#!/usr/bin/env python3
"""Per-segment share of the warp-time of ONE captured kernel launch (the first in the report): the SASS is cut at every
BAR / WARPSYNC.ALL, and the samples stalled on `barrier` (warps waiting for the slowest role) are shown apart from the
samples of warps that are busy in the segment.      python tools/ncu_segments.py gpurun_out/x.ncu-rep > profiles/x_phases.txt"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_idx = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
h = rows[hdr_idx[0]]
ix = {k: i for i, k in enumerate(h)}
data = rows[hdr_idx[0] + 1:(hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows))]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
def new(): return {"n": 0, "s": 0.0, "bar": 0.0, "st": {}, "inst": 0.0}
segs, cur = [], new()
for r in data:
    src = r[ix["Source"]]
    cur["n"] += 1; cur["s"] += f(r, "# Samples"); cur["bar"] += f(r, "stall_barrier"); cur["inst"] += f(r, "Instructions Executed")
    for k in stalls: cur["st"][k] = cur["st"].get(k, 0) + f(r, k)
    if "BAR." in src or "WARPSYNC.ALL" in src:
        segs.append(cur); cur = new()
segs.append(cur)
tot = sum(s["s"] for s in segs) or 1
print("# %s: SASS segments between barriers of the first captured launch; share of all stall samples (= warp-time)," % sys.argv[1])
print("# split into warps waiting at the barrier that opens the segment and warps busy in it; executed warp instructions; top stalls")
for i, sg in enumerate(segs):
    if sg["s"] < 0.004 * tot: continue
    top = sorted(sg["st"].items(), key=lambda kv: -kv[1])[:4]
    print("seg%-2d static=%5d share=%5.1f%% (barrier-wait %4.1f%%, busy %4.1f%%) inst=%6.1fM  %s" % (
        i, sg["n"], 100 * sg["s"] / tot, 100 * sg["bar"] / tot, 100 * (sg["s"] - sg["bar"]) / tot, sg["inst"] / 1e6,
        " ".join("%s=%.0f%%" % (k[6:], 100 * v / sg["s"]) for k, v in top)))
