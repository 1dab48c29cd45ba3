#!/usr/bin/env python3
"""Offline study of tree quality (r02): the SAH cost of the GPU-built LBVH (tools/dump_bvh.py) against a full-sweep SAH tree
over the SAME leaves — the bound of what a tree-restructuring pass could gain in node visits.
usage: sah_study.py gpurun_out/bvh_cow.npz ..."""
import sys
import numpy as np
sys.setrecursionlimit(100000)

def area(lo, hi):
    d = np.maximum(hi - lo, 0.0)
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]

def lbvh_cost(nodes):
    """sum of SA over internal nodes (= expected pair visits up to a constant) and leaves collected"""
    leaves_lo, leaves_hi, leaves_n = [], [], []
    total = 0.0
    stack = [(0, None, None)]
    # root box = union of pair 0's children
    while stack:
        p, lo, hi = stack.pop()
        a, b = nodes[2 * p], nodes[2 * p + 1]
        recs = [r for r in (a, b) if not (np.isinf(r["bmin"]).any() and r["meta"] == 0)]
        blo = np.min([r["bmin"] for r in recs], axis=0); bhi = np.max([r["bmax"] for r in recs], axis=0)
        total += area(blo, bhi)
        for r in recs:
            if r["link"] >= 0:
                stack.append((int(r["link"]), r["bmin"], r["bmax"]))
            else:
                leaves_lo.append(r["bmin"]); leaves_hi.append(r["bmax"]); leaves_n.append(int(r["meta"]))
    return total, np.array(leaves_lo, np.float64), np.array(leaves_hi, np.float64), np.array(leaves_n)

def sweep_sah(lo, hi, idx):
    """full-sweep SAH over leaves idx: returns sum of SA over internal nodes"""
    if len(idx) == 1:
        return 0.0
    blo = lo[idx].min(axis=0); bhi = hi[idx].max(axis=0)
    me = area(blo, bhi)
    if len(idx) == 2:
        return me
    best = None
    cen = 0.5 * (lo[idx] + hi[idx])
    for ax in range(3):
        order = idx[np.argsort(cen[:, ax], kind="stable")]
        l_lo = np.minimum.accumulate(lo[order], axis=0); l_hi = np.maximum.accumulate(hi[order], axis=0)
        r_lo = np.minimum.accumulate(lo[order][::-1], axis=0)[::-1]; r_hi = np.maximum.accumulate(hi[order][::-1], axis=0)[::-1]
        n = len(order)
        k = np.arange(1, n)
        cost = area(l_lo[:-1], l_hi[:-1]) * k + area(r_lo[1:], r_hi[1:]) * (n - k)
        j = int(np.argmin(cost))
        if best is None or cost[j] < best[0]:
            best = (cost[j], order[: j + 1], order[j + 1:])
    return me + sweep_sah(lo, hi, best[1]) + sweep_sah(lo, hi, best[2])

for path in sys.argv[1:]:
    d = np.load(path)
    nodes = d["nodes"]
    c_lbvh, llo, lhi, ln = lbvh_cost(nodes)
    root = area(llo.min(axis=0), lhi.max(axis=0))
    if len(ln) > 60000:
        print(path, "leaves", len(ln), "LBVH internal SA / root", c_lbvh / root, "(too large for the sweep build)")
        continue
    c_sah = sweep_sah(llo, lhi, np.arange(len(ln)))
    print(f"{path}: {len(ln)} leaves ({ln.sum()} prims)  sum SA(internal)/SA(root): LBVH {c_lbvh / root:.2f}  sweep-SAH {c_sah / root:.2f}  ratio {c_lbvh / c_sah:.2f}")
