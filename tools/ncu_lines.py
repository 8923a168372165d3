#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep (needs -lineinfo + --import-source on).
   python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
import os
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = os.environ.get("NCU_FILE")   # restrict to one source file, e.g. NCU_FILE=preprocess.cu
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; acc = collections.OrderedDict(); fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[0].strip():
        if only and fname != only: continue
        key = (fname, r[0], r[1].strip())
        a = acc.setdefault(key, [0.0, 0.0, 0])
        def num(x):
            try: return float(x)
            except ValueError: return 0.0
        a[0] += num(r[ii]); a[1] += num(r[sm]); a[2] += 1
ti = sum(v[0] for v in acc.values()); ts = sum(v[1] for v in acc.values())
print(f"total warp-instructions {ti:.4g}, samples {ts:.4g}")
# optional region roll-up: tools/ncu_lines.py rep N "name:lo-hi,name:lo-hi"
if len(sys.argv) > 3:
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi = [int(x) for x in rng.split("-")]
        vi = sum(v[0] for (f, ln, s_), v in acc.items() if lo <= int(ln) <= hi)
        vs = sum(v[1] for (f, ln, s_), v in acc.items() if lo <= int(ln) <= hi)
        print(f"region {name:12s} lines {lo}-{hi}: inst={vi/ti*100:5.1f}% samples={vs/ts*100:5.1f}%")
for (f, ln, src), v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f}:{ln:>4} inst={v[0]/ti*100:5.1f}% samples={v[1]/ts*100:5.1f}% sass={v[2]:3d} | {src[:100]}")
