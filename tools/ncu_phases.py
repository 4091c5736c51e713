#!/usr/bin/env python3
"""Per-phase stall breakdown of a solve kernel from an ncu report: the SASS is cut at every barrier / warp-sync and
the source-page samples are summed per segment.   python tools/ncu_phases.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs, cur = [], None
def new(): return {"n": 0, "samples": 0.0, "inst": 0.0, "thr": 0.0, "st": {}}
cur = new()
for r in data:
    src = r[ix["Source"]]
    cur["n"] += 1; cur["samples"] += f(r, "# Samples"); cur["inst"] += f(r, "Instructions Executed"); cur["thr"] += f(r, "Thread Instructions Executed")
    for s in stalls: cur["st"][s] = cur["st"].get(s, 0) + f(r, s)
    if "BAR." in src or "WARPSYNC" in src or "EXIT" in src:
        cur["end"] = src.strip()[:34]; segs.append(cur); cur = new()
segs.append(cur)
tot = sum(s["samples"] for s in segs) or 1
print("# %s: segments of the kernel's SASS between barriers; share of stall samples, executed warp instructions" % sys.argv[1])
for i, s in enumerate(segs):
    if s["samples"] < 0.003 * tot: continue
    top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:6]
    print("seg%-2d static=%5d samples=%5.1f%% inst=%7.1fM lanes=%4.1f end=[%s]  %s" % (i, s["n"], 100 * s["samples"] / tot, s["inst"] / 1e6, s["thr"] / max(s["inst"], 1), s.get("end", ""), " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(s["samples"], 1)) for k, v in top)))
