#!/usr/bin/env python3
"""Per-CUDA-source-line profile of one kernel from an .ncu-rep captured with --import-source on.

usage: ncu_src.py <file.ncu-rep> <kernel regex> [top N]
Prints the source lines with the most executed warp instructions / stall samples (ncu --page source --print-source cuda,sass)."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1:3]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fpath, hdr, recs = None, None, []
    seen_fn = set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            key = (fpath, fn)
            skip = key in seen_fn  # one launch only: the first instance of (file, kernel)
            seen_fn.add(key)
            continue
        if hdr is None or skip or len(r) < 10 or r[2] != "-":
            continue  # r[2] == '-' marks an aggregated CUDA line; SASS rows carry an address
        try:
            i_s, i_ie, i_te = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            recs.append((fpath, int(r[0]), r[1].strip(), int(r[i_s] or 0), int(r[i_ie] or 0), int(r[i_te] or 0)))
        except ValueError:
            pass
    ts = sum(x[3] for x in recs) or 1
    ti = sum(x[4] for x in recs) or 1
    tt = sum(x[5] for x in recs)
    print(f"kernel {kern}: samples {ts}  warp-inst {ti}  thread-inst {tt}  lanes/inst {tt / ti:.2f}")
    print("  inst%  samp%  lanes  file:line  source")
    for f, l, s, ns, ie, te in sorted(recs, key=lambda x: -x[3 if "--by-samples" in sys.argv else 4])[:top]:
        print(f"  {100 * ie / ti:5.1f}  {100 * ns / ts:5.1f}  {te / max(ie, 1):5.1f}  {f}:{l}  {s[:110]}")


if __name__ == "__main__":
    main()
