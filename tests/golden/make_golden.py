#!/usr/bin/env python3
"""Write tests/golden/oracle_hits_v1.npz: ray batches and the oracle's closest hits for three scenes.

The reference ships no golden vectors and cannot be run here (Rust nightly), so this
file pins the ORACLE itself: it was generated once, by this script, with the oracle as first
committed and reviewed against the Rust sources; test_oracle_kat.py::test_oracle_matches_committed_golden
fails if a later edit changes any id / t / normal.   Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import Oracle  # noqa: E402
import raytracer_weekend_b200 as rtw  # noqa: E402


def main():
    orc = Oracle()
    out = {}
    for scene, aspect in (("cornell-box", 1.0), ("cow-lambert-metal", 16 / 9), ("jumpy-balls", 16 / 9)):
        w = 64
        h = int(round(w / aspect))
        with rtw.Scene.from_name(orc, scene, aspect, seed=1) as s:
            rays = np.concatenate([orc.capture_rays(s, s.cameras[0], w, h, 5, 0, b) for b in (0, 1, 2)])
            hits = s.trace_closest(rays)
        out[f"{scene}/rays"] = rays.view(np.uint8)
        out[f"{scene}/hits"] = hits.view(np.uint8)
        print(scene, len(rays), "rays,", int(np.count_nonzero(hits["prim_id"] >= 0)), "hits")
    np.savez_compressed(os.path.join(HERE, "oracle_hits_v1.npz"), **out)


if __name__ == "__main__":
    main()
