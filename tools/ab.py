#!/usr/bin/env python3
"""A/B harness for kernel variants (run on the GPU box).

    tools/ab.py --variants "base:" "x:-DRTW_FOO=1 -DRTW_BAR" --work cornell-box:200 cow:32

Every variant is librtw_cuda.so rebuilt with the given extra nvcc flags into lib/ab_<name>/ (build in
build_<name>/, done by tools/ab_build.sh here, before gpurun).  For each (variant, workload) it prints the
product path's Mrays/s (best of 3 frames after 2 warm-ups) and the per-kernel CUDA-event times of the
instrumented single-pool path (RTW_RENDER_TIME_KERNELS)."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import torch
import raytracer_weekend_b200 as rtw
from raytracer_weekend_b200 import api
api.CUDA_LIB = %(lib)r
gpu = rtw.cuda_backend()
import bench
name, spp = %(work)r.split(':'); spp = int(spp)
scene_name, w, h, _ = bench.WORKLOADS[name]
dev = torch.device('cuda', 0)
accum = torch.zeros(h * w * 3, device=dev, dtype=torch.float32)
stream = torch.cuda.current_stream(dev)
scene = rtw.Scene.from_name(gpu, scene_name, w / h, seed=2024, device=0)
cam = scene.cameras[0]
best = 0.0
for i in range(5):
    p = scene.params(w, h, spp, seed=2024, pool_size=int(os.environ.get('AB_POOL', '0')))
    st = scene.render_device(cam, p, accum.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    if i >= 2: best = max(best, st.segments / st.ms_render / 1e3)
p = scene.params(w, h, spp, seed=2024, flags=2, pool_size=int(os.environ.get('AB_POOL', '0')))
for i in range(2):
    st = scene.render_device(cam, p, accum.data_ptr(), stream.cuda_stream)
torch.cuda.synchronize()
print('%%-10s %%-16s %%8.1f Mrays/s | single pool: traverse %%8.2f ms  sort %%6.2f ms  shade %%8.2f ms  (%%d iterations, %%.1f Mseg, pool %%d, build %%.2f ms, height %%d)' %% (
    %(var)r, %(work)r, best, st.ms_traverse, st.ms_sort, st.ms_shade, st.iterations, st.segments / 1e6, st.pool_size, scene.build_stats.ms_build, scene.build_stats.max_depth))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", nargs="+", default=["base"])
    ap.add_argument("--work", nargs="+", default=["cornell-box:200", "cow:32"])
    a = ap.parse_args()
    for work in a.work:
        for v in a.variants:
            name = v.split(":")[0]
            lib = os.path.join(ROOT, "raytracer-weekend_b200", "lib", f"ab_{name}", "librtw_cuda.so")
            if name == "base" and not os.path.exists(lib):
                lib = os.path.join(ROOT, "raytracer-weekend_b200", "lib", "librtw_cuda.so")
            code = CHILD % {"root": ROOT, "lib": lib, "work": work, "var": name}
            r = subprocess.run([sys.executable, "-c", code], text=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            sys.stdout.write(r.stdout if r.returncode == 0 else f"{name} {work} FAILED: {r.stderr[-400:]}\n")
            sys.stdout.flush()


if __name__ == "__main__":
    main()
