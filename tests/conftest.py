"""Shared fixtures.  GPU tests are marked `gpu` and call the product through its C ABI; the CPU
oracle (oracle/liboracle.so) is loaded HERE and only here — it is the checker, never the product."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import raytracer_weekend_b200 as rtw  # noqa: E402

ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """Missing libraries: the full build().  Present: `make` in the three directories — a no-op when everything is up to date
    (0.2 s), a rebuild when a source is newer than its library, so the suite never tests a stale binary."""
    import __graft_entry__ as g

    need = [rtw.CUDA_LIB, rtw.HOST_LIB, ORACLE_LIB]
    if not all(os.path.exists(p) for p in need):
        g.build()
        return
    if os.environ.get("RTW_CUDA_LIB"):   # an A/B build was asked for explicitly: leave it alone
        return
    if os.path.exists("/dev/nvidiactl"):  # a GPU box: the libraries travelled with the snapshot (built and tested here), file
        return                            # times may not have — do not spend the box's time on a rebuild of the same code
    for d in (os.path.join(g.PKG, "csrc"), os.path.join(g.PKG, "host"), os.path.join(ROOT, "oracle")):
        try:
            g._make(d)
        except Exception as e:   # a box without the toolchain: test the libraries that travelled with the snapshot
            sys.stderr.write(f"conftest: could not refresh {d} ({e}); using the existing libraries\n")


_ensure_built()


class Oracle(rtw.Backend):
    """The oracle's C API = the rtw_ ABI under the prefix orc_ plus checker-only helpers."""

    def __init__(self):
        super().__init__(ORACLE_LIB, "orc_")
        f = self.fn
        f("render_ex").argtypes = [C.c_void_p, C.POINTER(rtw.Camera), C.POINTER(rtw.RenderParams), C.c_void_p,
                                   C.POINTER(rtw.RenderStats), C.c_int, C.c_int, C.c_int]
        f("capture_rays").argtypes = [C.c_void_p, C.POINTER(rtw.Camera), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                      C.c_uint32, C.c_void_p]
        f("philox4x32_10").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("philox4x32_10").restype = None
        f("rng_draws").argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_uint32,
                                   C.c_void_p, C.c_void_p]
        f("texture_value").argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        f("aabb_hit").argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
        f("camera_new").argtypes = [C.POINTER(C.c_float)] * 3 + [C.c_float] * 6 + [C.POINTER(rtw.Camera)]

    def render_ex(self, scene, cam, params, mode=0, integrator=0, threads=0):
        accum = np.zeros((params.height, params.width, 3), np.float32)
        st = rtw.RenderStats()
        self.check(self.fn("render_ex")(scene.h, C.byref(cam), C.byref(params), accum.ctypes.data, C.byref(st), mode,
                                        integrator, threads), "render_ex")
        return accum, st

    def capture_rays(self, scene, cam, w, h, seed, sample, bounce):
        rays = np.zeros(w * h, rtw.RAY_DTYPE)
        self.check(self.fn("capture_rays")(scene.h, C.byref(cam), w, h, seed, sample, bounce, rays.ctypes.data), "capture_rays")
        return rays

    def camera_new(self, look_from, look_at, up, vfov, aspect, aperture, focus, t0=0.0, t1=1.0):
        cam = rtw.Camera()
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])  # noqa: E731
        self.fn("camera_new")(f3(look_from), f3(look_at), f3(up), vfov, aspect, aperture, focus, t0, t1, C.byref(cam))
        return cam

    def texture_value(self, scene, tex, u, v, p):
        out = (C.c_float * 3)()
        self.check(self.fn("texture_value")(scene.h, tex, u, v, (C.c_float * 3)(*[float(x) for x in p]), out), "texture_value")
        return np.array(out[:], np.float32)


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


def _gpu_available():
    try:
        return rtw.cuda_backend().device_count() > 0
    except Exception:
        return False


HAS_GPU = _gpu_available()


@pytest.fixture(scope="session")
def gpu():
    if not HAS_GPU:
        pytest.skip("no CUDA device")
    return rtw.cuda_backend()


def pytest_collection_modifyitems(config, items):
    # a `gpu` test on a box without a GPU is an error of the invocation, not a pass: skip loudly
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
