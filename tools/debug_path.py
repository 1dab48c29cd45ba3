"""Find why the LBVH traversal misses a hit the brute force finds: walk root->leaf path of the missed primitive
and evaluate the slab test (float32, same op order as rtw_traverse.cuh::slab) at every record."""
import sys
import numpy as np
sys.path.insert(0, ".")
import raytracer_weekend_b200 as rtw

name = sys.argv[1]
n = int(sys.argv[2])
gpu = rtw.cuda_backend()
s = rtw.Scene.from_name(gpu, name, 16 / 9, seed=2024)
rs = np.random.RandomState(0)
o = np.tile([[0, 0, -260]], (n, 1)).astype(np.float32)
d = (np.array([[0, 0, 1]]) + rs.uniform(-.3, .3, (n, 3)) * [1, 1, 0]).astype(np.float32)
rays = rtw.make_rays(o, d)
hb = s.trace_closest(rays, rtw.RTW_TRACE_BRUTE)
hv = s.trace_closest(rays, rtw.RTW_TRACE_BVH)
bad = np.nonzero(hb["prim_id"] != hv["prim_id"])[0]
print("mismatches", len(bad))
nodes, slot_ids, root = s.get_bvh()
npairs = len(nodes) // 2
# parents: for every internal child link
parent = np.full(npairs, -1, np.int64)
inner = nodes["link"] >= 0
parent[nodes["link"][inner]] = np.nonzero(inner)[0] // 2
slot_of = np.zeros(len(slot_ids), np.int64); slot_of[slot_ids] = np.arange(len(slot_ids))
leaf_recs = np.nonzero((nodes["link"] < 0) & ~np.isinf(nodes["bmin"]).any(axis=1))[0]
first = ~nodes["link"][leaf_recs].astype(np.int64)
order = np.argsort(first)
leaf_recs, first = leaf_recs[order], first[order]
F = np.float32
def slab(rec, o, inv, tmin, tmax):
    for a in range(3):
        t0 = F(F(rec["bmin"][a] - o[a]) * inv[a]); t1 = F(F(rec["bmax"][a] - o[a]) * inv[a])
        tn, tf = (t1, t0) if inv[a] < 0 else (t0, t1)
        tmin = np.fmax(tn, tmin); tmax = np.fmin(tf, tmax)
    tfar = F(abs(tmax)) * F(4.76837158e-7) + tmax
    return tmin, tmax, tmin <= tfar
for i in bad[:4]:
    pid = int(hb["prim_id"][i]); slot = slot_of[pid]
    k = np.searchsorted(first, slot, side="right") - 1
    rec = leaf_recs[k]
    print(f"ray {i}: prim {pid} slot {slot} t {hb['t'][i]} leaf rec {rec} range [{first[k]}, +{nodes['meta'][rec]}) ; bvh found {hv['prim_id'][i]} t {hv['t'][i]}")
    inv = (F(1) / d[i]).astype(np.float32)
    chain = [rec]
    p = rec // 2
    while p != 0:
        pp = parent[p]
        side = 0 if nodes["link"][2 * pp] == p else 1
        chain.append(2 * pp + side)
        p = pp
    for r in reversed(chain):
        tmin, tmax, ok = slab(nodes[r], o[i], inv, F(0.001), F(np.inf))
        print(f"   rec {r} link {nodes['link'][r]} meta {nodes['meta'][r]} box {nodes['bmin'][r]} {nodes['bmax'][r]} -> tmin {tmin} tmax {tmax} pass {ok}")
    bc = (nodes["bmin"][rec] + nodes["bmax"][rec]) / 2; br = (nodes["bmax"][rec] - nodes["bmin"][rec]) / 2
    print("   leaf box centre", bc, "half", br, " hit p", hb["p"][i], "n", hb["normal"][i], " centre from hit with r=half.x:", hb["p"][i] - hb["normal"][i] * br[0],
          "dist p-boxcentre", np.linalg.norm(hb["p"][i] - bc))
