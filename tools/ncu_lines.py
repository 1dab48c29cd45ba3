#!/usr/bin/env python3
"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info: samples / instructions per source line.

usage: ncu_lines.py <ncu_source.csv> <nvdisasm -g -c output> <kernel substring> [top N]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, sass_path, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # nvdisasm: sections per function; lines "//## File "x", line N" precede instructions "/*0010*/ ..."
    addr2line = {}
    in_fn = False
    cur = None
    inl = ""
    for line in open(sass_path, errors="replace"):
        if line.startswith(".text."):
            in_fn = kernel in line
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            inl = m.group(3)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r)
    h = rows[hi]
    ai, ns, ie, te = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    base = None
    agg = defaultdict(lambda: [0, 0, 0])
    for r in rows[hi + 1:]:
        if len(r) <= te or not r[ai]:
            continue
        if r[ai] == "Address":  # next kernel section
            break
        a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
        if base is None:
            base = a
        key = addr2line.get(a - base, ("?", 0))
        agg[key][0] += int(r[ns] or 0)
        agg[key][1] += int(r[ie] or 0)
        agg[key][2] += int(r[te] or 0)
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    tot_t = sum(v[2] for v in agg.values())
    print(f"samples {tot_s}  warp-inst {tot_i}  thread-inst {tot_t}  avg threads/inst {tot_t / tot_i:.2f}")
    srcs = {}
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            try:
                srcs[f] = open(f"/root/repo/raytracer-weekend_b200/csrc/{f}").read().split("\n")
            except OSError:
                srcs[f] = []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print(f"{v[0] / tot_s * 100:5.1f}% smp {v[1] / tot_i * 100:5.1f}% inst  thr/inst {v[2] / max(v[1], 1):5.1f}  {f}:{l:<4d} {text}")


if __name__ == "__main__":
    main()
