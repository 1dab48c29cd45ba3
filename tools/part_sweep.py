#!/usr/bin/env python3
"""Strong-scaling emulation on one GPU: render rank 0's share of an N-way tile partition for several (pool, slices).
    tools/part_sweep.py <workload> <spp> <parts> pool:slices [pool:slices ...]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raytracer_weekend_b200 as rtw
import bench
name, spp, parts = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
scene_name, w, h, _ = bench.WORKLOADS[name]
gpu = rtw.cuda_backend()
dev = torch.device('cuda', 0)
accum = torch.zeros(h * w * 3, device=dev, dtype=torch.float32)
stream = torch.cuda.current_stream(dev)
scene = rtw.Scene.from_name(gpu, scene_name, w / h, seed=2024, device=0)
cam = scene.cameras[0]
for spec in sys.argv[4:]:
    pool, slices = (int(x) for x in spec.split(':'))
    best = None
    for i in range(4):
        p = scene.params(w, h, spp, seed=2024, pool_size=pool, slices=slices, part_rank=0, part_count=parts)
        st = scene.render_device(cam, p, accum.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        if i >= 1 and (best is None or st.ms_render < best.ms_render):
            best = st
    print('%s 1/%d of the tiles: pool %8d slices %4d -> %8.2f ms  %7.1f Mrays/s  %d iterations' % (
        name, parts, best.pool_size, best.slices, best.ms_render, best.segments / best.ms_render / 1e3, best.iterations))
