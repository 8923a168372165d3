#!/usr/bin/env python
"""DRAM traffic per step from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch list of `bench.py --steps 2 --warmup 1`: a step starts at preprocess_plan_kernel; the third step (second timed
step of the device-resident loop) is summarised into profiles/ncu_traffic.json, which bench.py copies into
`roofline.traffic` / `roofline_preprocess.traffic`.   python tools/ncu_traffic.py launches.csv out.json"""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
ki, mi, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
launches = {}
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr):
        continue
    d = launches.setdefault(int(r[idi]), {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
steps, cur = [], None
for i in sorted(launches):
    l = launches[i]
    if "preprocess_plan_kernel" in l["name"]:
        cur = []
        steps.append(cur)
    if cur is not None:
        cur.append(l)
step = steps[2]
def agg(pred):
    sel = [l for l in step if pred(l["name"])]
    return {"launches": len(sel), "dram_bytes": sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in sel),
            "time_us": sum(l.get("gpu__time_duration.sum", 0) for l in sel)}
out = {"source": sys.argv[1], "step_index": 2, "kernels_in_step": len(step),
       "conv": agg(lambda n: "conv" in n), "preprocess": agg(lambda n: "preprocess_kernel(" in n or "preprocess_tc_kernel(" in n),
       "all": agg(lambda n: True),
       "per_kernel": [{"name": l["name"][:60], "dram_bytes": l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0),
                       "time_us": l.get("gpu__time_duration.sum", 0)} for l in step]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print({k: out[k] for k in ("kernels_in_step", "conv", "preprocess", "all")})
