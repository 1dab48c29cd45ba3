#!/usr/bin/env python3
"""animated-book2-final-scene (scenes.rs:622-667, 30 cameras over one world): rtw_render_frames with a callback that
tonemaps + PNG-encodes every frame (what main.rs:66-94 does per frame) vs the same work done frame by frame with
rtw_render.  Shows the overlap of D2H + host work of frame n with the rendering of frame n+1.
    tools/anim_bench.py [width] [spp] [frames]"""
import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from PIL import Image
import raytracer_weekend_b200 as rtw

w = int(sys.argv[1]) if len(sys.argv) > 1 else 800
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 100
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 30
gpu = rtw.cuda_backend()
t0 = time.perf_counter()
scene = rtw.Scene.from_name(gpu, "animated-book2-final-scene", 1.0, seed=1)
t_build = time.perf_counter() - t0
cams = scene.cameras[:nf]
p = scene.params(w, w, spp, seed=7)


def encode(accum):
    c = np.sqrt(accum / spp)
    img = (255.999 * np.clip(c, 0, 0.999)).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="PNG")
    return buf.tell()


seg = [0]
host = [0.0]


def on_frame(i, accum, st):
    t = time.perf_counter()
    encode(accum)
    host[0] += time.perf_counter() - t
    seg[0] += st["segments"]


scene.render_frames(cams[:2], p, lambda i, a, st: None)  # warm up
t0 = time.perf_counter()
n = scene.render_frames(cams, p, on_frame)
t_overlap = time.perf_counter() - t0
host_overlap = host[0]
t0 = time.perf_counter()
gpu_ms = 0.0
for i, cam in enumerate(cams):
    q = scene.params(w, w, spp, seed=7 + i)
    a, st = scene.render(cam, q)
    gpu_ms += st.ms_render
    encode(a)
t_serial = time.perf_counter() - t0
print(f"animated-book2-final-scene {w}x{w} {spp} spp, {n} frames, scene build {t_build:.2f} s (once)")
print(f"  rtw_render_frames + PNG encode in the callback : {t_overlap:.3f} s  ({seg[0] / t_overlap / 1e6:.0f} Mrays/s end to end; host encode {host_overlap:.3f} s hidden)")
print(f"  rtw_render per frame, then PNG encode (serial)  : {t_serial:.3f} s  (GPU render {gpu_ms / 1e3:.3f} s)")
