"""raytracer-weekend_b200 — B200-native path-tracing backend for raytracer_weekend_lib.

Layout:
  csrc/   hand-written CUDA (sm_100a) + the C ABI of include/rtw_cuda.h      -> lib/librtw_cuda.so
  host/   C++ mirror of the reference's scene API (flatten), scenes, console_app -> lib/librtw_host.so, bin/console_app
  api.py  ctypes plumbing for the Python harness (tests, bench.py)

The directory name contains a hyphen; import it as ``raytracer_weekend_b200`` (a shim package at the
repository root extends its __path__ to this directory).
"""
from .api import *  # noqa: F401,F403
from .api import Backend, Scene, cuda_backend, host_lib  # noqa: F401
