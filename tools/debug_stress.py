import sys, time
import numpy as np
sys.path.insert(0, ".")
import raytracer_weekend_b200 as rtw
name = sys.argv[1] if len(sys.argv) > 1 else "stress:1000000:1000000"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
gpu = rtw.cuda_backend()
s = rtw.Scene.from_name(gpu, name, 16/9, seed=2024)
rs = np.random.RandomState(0)
o = np.tile([[0, 0, -260]], (n, 1)).astype(np.float32)
d = (np.array([[0, 0, 1]]) + rs.uniform(-.3, .3, (n, 3)) * [1, 1, 0]).astype(np.float32)
rays = rtw.make_rays(o, d)
hb = s.trace_closest(rays, rtw.RTW_TRACE_BRUTE)
hv = s.trace_closest(rays, rtw.RTW_TRACE_BVH)
bad = np.nonzero((hb["prim_id"] != hv["prim_id"]) | (hb["t"].view(np.uint32) != hv["t"].view(np.uint32)))[0]
print("mismatches", len(bad), "of", n)
for i in bad[:10]:
    print(i, "brute", hb["prim_id"][i], hb["t"][i], "bvh", hv["prim_id"][i], hv["t"][i], "types", s.prim_info(int(hb["prim_id"][i])) if hb["prim_id"][i] >= 0 else None,
          s.prim_info(int(hv["prim_id"][i])) if hv["prim_id"][i] >= 0 else None)
# (the CPU oracle is test infrastructure: to see what it says about these rays use tests/, e.g.
#  test_large_scene_mismatches_are_only_ill_conditioned_reference_hits)
