import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import raytracer_weekend_b200 as rtw
gpu = rtw.cuda_backend()
torch.zeros(1, device='cuda')
for name, w, h, spp in (("jumpy-balls", 400, 225, 100), ("cornell-box", 800, 800, 100), ("cow-lambert-metal", 1920, 1080, 16)):
    world = rtw.World(name, w / h, seed=2024)
    host = np.zeros((h, w, 3), np.float32)
    for rep in range(3):
        t0 = time.perf_counter()
        s = rtw.Scene.from_world(gpu, world, device=0)
        t1 = time.perf_counter()
        p = s.params(w, h, spp, seed=1)
        _, st = s.render(s.cameras[0], p, out=host)
        t2 = time.perf_counter()
        _, st2 = s.render(s.cameras[0], p, out=host)
        t3 = time.perf_counter()
        bs = s.build_stats
        s.close()
        t4 = time.perf_counter()
        print(f"{name:18s} rep {rep}: flatten+build {1e3*(t1-t0):7.2f} ms (gpu build {bs.ms_build:.2f}, upload {bs.ms_upload:.2f})  first render call {1e3*(t2-t1):7.2f} (device {st.ms_render:.2f})  second {1e3*(t3-t2):7.2f} (device {st2.ms_render:.2f})  close {1e3*(t4-t3):.2f}")
    world.close()
