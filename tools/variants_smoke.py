#!/usr/bin/env python3
"""Small renders / traces through every kernel variant (general, compact, 4-wide, flat, media, counting, timing,
frames): a crash / hang canary.  (Written for `compute-sanitizer --tool memcheck`, which is closed on the r01 pool.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_weekend_b200 as rtw

gpu = rtw.cuda_backend()
for env in ({}, {"RTW_COMPACT": "1"}, {"RTW_WIDE": "1"}, {"RTW_FLAT": "0"}):
    os.environ.update(env)
    for name, aspect in (("cornell-box", 1.0), ("cow-lambert-metal", 16 / 9), ("smokey-cornell-box", 1.0), ("stress:2000:300", 16 / 9)):
        with rtw.Scene.from_name(gpu, name, aspect, seed=1) as s:
            cam = s.cameras[0]
            for (w, h, spp, pool, slices) in ((40, 24, 3, 0, 0), (33, 17, 2, 64, 2)):
                a, st = s.render(cam, s.params(w, h, spp, seed=3, pool_size=pool, slices=slices))
                assert np.isfinite(a).all() and st.paths == w * h * spp
            s.render(cam, s.params(32, 32, 2, seed=3, flags=rtw.RTW_RENDER_COUNT_TRAVERSAL | rtw.RTW_RENDER_TIME_KERNELS))
            rays = rtw.make_rays(np.tile([[0.0, 1.0, -5.0]], (257, 1)), np.random.RandomState(0).uniform(-1, 1, (257, 3)))
            s.trace_closest(rays)
            s.trace_closest(rays, mode=rtw.RTW_TRACE_BRUTE)
            s.render_frames([cam, cam], s.params(16, 16, 1, seed=1), lambda i, a, st: True)
    for k in env:
        del os.environ[k]
print("variants_smoke: all variants ran")
