"""Vectorised structural check of a GPU-built LBVH: coverage (every slot in exactly one leaf), containment."""
import sys
import numpy as np
sys.path.insert(0, ".")
import raytracer_weekend_b200 as rtw


def check(scene):
    nodes, slot_ids, root = scene.get_bvh()
    n = scene.num_prims
    assert np.array_equal(np.sort(slot_ids), np.arange(n)), "slot_prim is not a permutation"
    cover = np.zeros(n + 1, np.int64)
    frontier = np.array([0], np.int64)
    plo = root[None, :3].copy(); phi = root[None, 3:].copy()
    visited = 0; leaves = 0; bad_contain = 0; depth = 0
    while len(frontier):
        depth += 1
        visited += len(frontier)
        recs = np.stack([nodes[2 * frontier], nodes[2 * frontier + 1]], axis=1)  # [m, 2]
        lo = np.repeat(plo, 2, axis=0); hi = np.repeat(phi, 2, axis=0)
        r = recs.reshape(-1)
        empty = np.isinf(r["bmin"]).any(axis=1) & (r["meta"] == 0)
        ok = (r["bmin"] >= lo - 1e-5 * (1 + np.abs(lo))).all(axis=1) & (r["bmax"] <= hi + 1e-5 * (1 + np.abs(hi))).all(axis=1)
        bad_contain += int(np.count_nonzero(~ok & ~empty))
        is_leaf = (r["link"] < 0) & ~empty
        first = (~r["link"][is_leaf]).astype(np.int64); cnt = r["meta"][is_leaf].astype(np.int64)
        np.add.at(cover, first, 1); np.add.at(cover, first + cnt, -1)
        leaves += int(is_leaf.sum())
        inner = r["link"] >= 0
        frontier = r["link"][inner].astype(np.int64)
        plo = r["bmin"][inner]; phi = r["bmax"][inner]
    c = np.cumsum(cover)[:n]
    return dict(n=n, visited_pairs=visited, leaves=leaves, depth=depth, uncovered=int((c == 0).sum()), multiply=int((c > 1).sum()),
                bad_contain=bad_contain)


if __name__ == "__main__":
    gpu = rtw.cuda_backend()
    for name in sys.argv[1:]:
        with rtw.Scene.from_name(gpu, name, 16 / 9, seed=2024) as s:
            print(name, check(s))
