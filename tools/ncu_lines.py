#!/usr/bin/env python3
"""Per-source-line profile of a kernel from an ncu report captured with --import-source on (code built with -lineinfo):
stall samples, executed warp instructions and shared-memory wavefronts per CUDA source line, grouped by file, top N.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
files, cur, hdr = {}, None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if cur is None or hdr is None or r[0] == "": continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def g(k):
        try:
            return float(r[hdr[k]] or 0) if k in hdr and r[hdr[k]] not in ("", "-") else 0.0
        except ValueError:   # (a column that lists access sizes, e.g. "32(18),64(72)")
            return 0.0
    d = files.setdefault(cur, {}).setdefault(line, {"src": r[1].strip(), "samples": 0.0, "inst": 0.0, "thr": 0.0, "wf": 0.0, "wfx": 0.0})
    d["samples"] += g("# Samples"); d["inst"] += g("Instructions Executed"); d["thr"] += g("Thread Instructions Executed")
    d["wf"] += g("L1 Wavefronts Shared"); d["wfx"] += g("L1 Wavefronts Shared Excessive")
tot_s = sum(d["samples"] for f in files.values() for d in f.values()) or 1
tot_i = sum(d["inst"] for f in files.values() for d in f.values()) or 1
print("# %s: total stall samples %.0f, executed warp instructions %.1f M" % (sys.argv[1], tot_s, tot_i / 1e6))
allrows = [(d["samples"], f, l, d) for f, ls in files.items() for l, d in ls.items()]
allrows.sort(key=lambda x: -x[0])
for s, f, l, d in allrows[:N]:
    print("%5.1f%% samples %5.1f%% inst lanes %4.1f smem wf %6.1fM (excess %5.1fM)  %s:%d  %s" % (100 * s / tot_s, 100 * d["inst"] / tot_i, d["thr"] / max(d["inst"], 1), d["wf"] / 1e6, d["wfx"] / 1e6, f.split("/")[-1], l, d["src"][:110]))
