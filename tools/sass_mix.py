#!/usr/bin/env python3
"""Static SASS summary per kernel of librtw_cuda.so: instruction count, local-memory ops, key opcodes.
usage: sass_mix.py [lib] [kernel substring ...]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else "raytracer-weekend_b200/lib/librtw_cuda.so"
subs = [a for a in sys.argv[1:] if not a.endswith(".so")] or ["k_wave", "k_mega"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, mix = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::|rtw::", "", fn)[:70]
        mix[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        mix[fn][m.group(1)] += 1
        mix[fn]["total"] += 1
KEYS = ["LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "ATOM", "RED", "MUFU", "FFMA", "FMUL", "FADD", "IMAD", "LOP3", "SHFL", "VOTE", "BSSY", "BRA", "UBLKCP", "UTMALDG", "SYNCS"]
for fn, c in mix.items():
    if not any(s in fn for s in subs):
        continue
    agg = collections.Counter()
    for op, n in c.items():
        for k in KEYS:
            if op.startswith(k):
                agg[k] += n
    print(f"{fn}\n   total {c['total']}: " + ", ".join(f"{k} {agg[k]}" for k in KEYS if agg[k]))
