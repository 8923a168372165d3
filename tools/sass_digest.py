#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (cuobjdump -sass of the shipped library):
UTC*MMA = tcgen05.mma (UTCHMMA half/bf16, UTCIMMA int8), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies,
UBLKCP = TMA bulk copies, UTCBAR = tcgen05.commit, HMMA would be the legacy mma.sync path.
    python tools/sass_digest.py [lib.so] > profiles/r02_sass_digest.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "playaid_core_b200", "libplayaid_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "IMMA", "SYNCS", "ACQBULK"]
cur, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter(); sizes[cur] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        sizes[cur] += 1
        op = m.group(1).split(".")[0]
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"SASS digest of {os.path.relpath(lib, ROOT)} (sm_100a): instructions per kernel and Blackwell mnemonics")
tot = collections.Counter()
for name, c in counts.items():
    tot.update(c)
    if sum(c.values()) == 0 and sizes[name] < 400:
        continue
    print(f"{name[:96]:96s} {sizes[name]:6d} instr  " + "  ".join(f"{k}={v}" for k, v in c.items() if v))
print("TOTAL  " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
