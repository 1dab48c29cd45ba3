#!/usr/bin/env python3
"""Dump the GPU-built LBVH of a scene (node records, slot order, root box) to gpurun_out/bvh_<tag>.npz for offline analysis
(tools/sah_study.py): usage dump_bvh.py scene[:tag] ..."""
import sys
import numpy as np
sys.path.insert(0, ".")
import raytracer_weekend_b200 as rtw
gpu = rtw.cuda_backend()
for arg in sys.argv[1:]:
    name, _, tag = arg.partition("=")
    with rtw.Scene.from_name(gpu, name, 16 / 9, seed=2024) as s:
        nodes, slots, root = s.get_bvh()
        np.savez_compressed(f"gpurun_out/bvh_{tag or name}.npz", nodes=nodes, slots=slots, root=root)
        print(name, len(nodes) // 2, "pairs", s.num_prims, "prims")
