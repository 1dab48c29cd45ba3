#!/usr/bin/env python3
"""Write tests/golden/oracle_renders_v1.npz: small frames rendered by the oracle (iterative integrator, flat world)
for scenes that together cover every material, texture, wrapper and the participating media.

Like oracle_hits_v1.npz this pins the ORACLE (the reference has no fixtures of its own): the file was
generated once by this script; tests/test_oracle_kat.py::test_oracle_matches_committed_render_golden fails if a later
edit changes a single bit of a frame, and the GPU suite compares the CUDA path with the same committed frames.
Usage: python tests/golden/make_golden_renders.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import Oracle  # noqa: E402
import raytracer_weekend_b200 as rtw  # noqa: E402

# scene, aspect, width, spp, slices
CASES = [("cornell-box", 1.0, 48, 6, 2), ("smokey-cornell-box", 1.0, 40, 4, 1), ("jumpy-balls", 16 / 9, 64, 4, 2),
         ("two-perlin-spheres", 16 / 9, 48, 3, 1), ("earth", 16 / 9, 48, 3, 1), ("simple-light", 16 / 9, 48, 4, 1),
         ("cow-lambert-metal", 16 / 9, 64, 3, 3)]
SEED = 4242


def main():
    orc = Oracle()
    out = {}
    for scene, aspect, w, spp, slices in CASES:
        h = int(round(w / aspect))
        with rtw.Scene.from_name(orc, scene, w / h, seed=1) as s:
            accum, st = s.render(s.cameras[0], s.params(w, h, spp, seed=SEED, slices=slices))
        out[f"{scene}/accum"] = accum
        out[f"{scene}/meta"] = np.array([w, h, spp, slices, st.segments], np.int64)
        print(scene, accum.shape, "segments", st.segments, "mean", float(accum.mean()))
    np.savez_compressed(os.path.join(HERE, "oracle_renders_v1.npz"), **out)


if __name__ == "__main__":
    main()
