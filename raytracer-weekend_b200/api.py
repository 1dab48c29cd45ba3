"""ctypes bindings of the C ABI (include/rtw_cuda.h, include/rtw_sink.h) and of the host front end.

This module is plumbing for the Python harness (pytest, bench.py, __graft_entry__): it loads
``lib/librtw_cuda.so`` (the CUDA backend) and ``lib/librtw_host.so`` (scene API mirror + scenes,
C++) and exposes thin numpy-friendly wrappers.  It contains no rendering logic of its own and no
CPU fallback: if the CUDA library is missing, loading fails loudly; if no GPU is present every
compute call raises :class:`RtwError` with the backend's message.

``Backend`` is generic over (shared library, symbol prefix) because the test-suite binds the CPU
oracle — which exports the same signatures with the prefix ``orc_`` — through the very same
wrapper.  Nothing in this package loads the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CUDA_LIB = os.environ.get("RTW_CUDA_LIB") or os.path.join(PKG_DIR, "lib", "librtw_cuda.so")  # env override: A/B builds
HOST_LIB = os.path.join(PKG_DIR, "lib", "librtw_host.so")
ASSET_DIR = os.path.join(REPO_ROOT, "assets")

RTW_TRACE_BVH = 0
RTW_TRACE_BRUTE = 1
RTW_RENDER_COUNT_TRAVERSAL = 1
RTW_RENDER_TIME_KERNELS = 2
RTW_OK, RTW_ERR_INVALID, RTW_ERR_CUDA, RTW_ERR_NOMEM, RTW_ERR_UNSUPPORTED, RTW_ERR_STATE = 0, -1, -2, -3, -4, -5

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("time", "<f4"), ("t_min", "<f4"), ("t_max", "<f4")])
HIT_DTYPE = np.dtype([("prim_id", "<i4"), ("material_id", "<i4"), ("t", "<f4"), ("p", "<f4", 3), ("normal", "<f4", 3),
                      ("u", "<f4"), ("v", "<f4"), ("front_face", "<i4")])
BVH_NODE_DTYPE = np.dtype([("bmin", "<f4", 3), ("link", "<i4"), ("bmax", "<f4", 3), ("meta", "<u4")])
assert RAY_DTYPE.itemsize == 36 and HIT_DTYPE.itemsize == 48 and BVH_NODE_DTYPE.itemsize == 32


class RtwError(RuntimeError):
    pass


class Camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("lens_radius", C.c_float), ("time0", C.c_float), ("time1", C.c_float)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("max_depth", C.c_uint32),
                ("background", C.c_float * 3), ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32),
                ("seed", C.c_uint64), ("tile_size", C.c_uint32), ("part_rank", C.c_uint32), ("part_count", C.c_uint32),
                ("pool_size", C.c_uint32), ("slices", C.c_uint32), ("flags", C.c_uint32), ("gpus", C.c_uint32),
                ("reserved", C.c_uint32)]


class RenderStats(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("paths", C.c_uint64), ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64),
                ("prim_bytes", C.c_uint64),
                ("iterations", C.c_uint32), ("launches", C.c_uint32), ("pool_size", C.c_uint32), ("slices", C.c_uint32),
                ("ms_render", C.c_float), ("ms_traverse", C.c_float), ("ms_shade", C.c_float), ("node_record_bytes", C.c_float),
                ("fused", C.c_uint32), ("gpus", C.c_uint32), ("ms_sort", C.c_float), ("ray_sort", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# rtw_frame_callback (include/rtw_cuda.h)
FRAME_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(RenderStats))


class BuildStats(C.Structure):
    _fields_ = [("num_prims", C.c_uint32), ("num_nodes", C.c_uint32), ("max_depth", C.c_uint32),
                ("num_instances", C.c_uint32), ("ms_build", C.c_float), ("ms_upload", C.c_float),
                ("device_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_SINK_FUNCS = ["last_error", "scene_create", "scene_destroy", "add_texture_solid", "add_texture_checker",
               "add_texture_noise", "add_texture_uvdebug", "add_texture_image", "add_material_lambertian",
               "add_material_metal", "add_material_dielectric", "add_material_diffuse_light", "push_translation",
               "push_rotation_y", "pop_transform", "begin_group", "end_group", "begin_medium", "end_medium", "add_sphere", "add_moving_sphere",
               "add_xy_rect", "add_xz_rect", "add_yz_rect", "add_cuboid", "add_triangles", "build", "render", "render_frames"]


class Sink(C.Structure):
    _fields_ = [("lib", C.c_void_p), ("scene", C.c_void_p)] + [(n, C.c_void_p) for n in _SINK_FUNCS]


def _fp(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _f3(v) -> C.Array:
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


class Backend:
    """One shared library exporting the rtw_cuda.h entry points under ``prefix``."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise RtwError(f"backend library not found: {path} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        self._setup()

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def has(self, name) -> bool:
        return hasattr(self.lib, self.prefix + name)

    def _setup(self):
        f = self.fn
        f("last_error").restype = C.c_char_p
        f("scene_create").argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        f("scene_destroy").argtypes = [C.c_void_p]
        f("add_texture_solid").argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        f("add_texture_checker").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        f("add_texture_noise").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        f("add_texture_uvdebug").argtypes = [C.c_void_p]
        f("add_texture_image").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        f("add_material_lambertian").argtypes = [C.c_void_p, C.c_int]
        f("add_material_metal").argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        f("add_material_dielectric").argtypes = [C.c_void_p, C.c_float]
        f("add_material_diffuse_light").argtypes = [C.c_void_p, C.c_int]
        f("push_translation").argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        f("push_rotation_y").argtypes = [C.c_void_p, C.c_float]
        f("push_rotation_y_sincos").argtypes = [C.c_void_p, C.c_float, C.c_float]
        f("pop_transform").argtypes = [C.c_void_p]
        f("begin_group").argtypes = [C.c_void_p]
        f("end_group").argtypes = [C.c_void_p]
        f("begin_medium").argtypes = [C.c_void_p, C.c_float, C.c_int]
        f("end_medium").argtypes = [C.c_void_p]
        f("add_sphere").argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_int]
        f("add_moving_sphere").argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float), C.c_float,
                                           C.c_float, C.c_int]
        for r in ("add_xy_rect", "add_xz_rect", "add_yz_rect"):
            f(r).argtypes = [C.c_void_p] + [C.c_float] * 5 + [C.c_int]
        f("add_cuboid").argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
        f("add_triangles").argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f("build").argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(BuildStats)]
        f("scene_num_prims").argtypes = [C.c_void_p]
        f("trace_closest").argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        f("render").argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(RenderParams), C.c_void_p, C.POINTER(RenderStats)]
        f("resolve_rgb8").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        f("render_frames").argtypes = [C.c_void_p, C.POINTER(Camera), C.c_uint32, C.POINTER(RenderParams), FRAME_CALLBACK,
                                       C.c_void_p]
        if self.has("render_device"):
            f("render_device").argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(RenderParams), C.c_void_p, C.c_void_p,
                                           C.POINTER(RenderStats)]
            f("trace_closest_device").argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
            f("get_bvh").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            f("scene_num_nodes").argtypes = [C.c_void_p]
            f("scene_num_instances").argtypes = [C.c_void_p]
            f("scene_prim_info").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
            f("scene_instance_ops").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        if self.has("scene_clone"):
            f("scene_clone").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]

    def last_error(self) -> str:
        return (self.fn("last_error")() or b"").decode(errors="replace")

    def check(self, rc: int, what: str) -> int:
        if rc < 0:
            raise RtwError(f"{self.prefix}{what} failed ({rc}): {self.last_error()}")
        return rc

    def device_count(self) -> int:
        return int(self.fn("device_count")())

    def live_handles(self) -> int:
        """rtw_debug_live_handles: CUDA events / streams / graphs the library holds right now."""
        return int(self.fn("debug_live_handles")())

    def trim_memory(self) -> int:
        """rtw_trim_memory: give the library's cache of freed device / pinned blocks back to the driver (MiB released)."""
        return int(self.fn("trim_memory")())

    def new_scene(self, device: int = 0) -> "Scene":
        return Scene(self, device)


_cuda_backend: Optional[Backend] = None
_host_lib = None


def cuda_backend() -> Backend:
    """The product backend.  Raises if librtw_cuda.so has not been built — there is no fallback."""
    global _cuda_backend
    if _cuda_backend is None:
        _cuda_backend = Backend(CUDA_LIB, "rtw_")
    return _cuda_backend


def host_lib():
    global _host_lib
    if _host_lib is None:
        if not os.path.exists(HOST_LIB):
            raise RtwError(f"host library not found: {HOST_LIB}")
        h = C.CDLL(HOST_LIB)
        h.rtwh_last_error.restype = C.c_char_p
        h.rtwh_capi_error.restype = C.c_char_p
        h.rtwh_sink_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(Sink)]
        h.rtwh_sink_close.argtypes = [C.POINTER(Sink)]
        h.rtwh_set_asset_dir.argtypes = [C.c_char_p]
        h.rtwh_register_image.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        h.rtwh_register_mesh.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        h.rtwh_scene_names.argtypes = [C.c_char_p, C.c_size_t]
        h.rtwh_build_scene.argtypes = [C.c_char_p, C.c_float, C.c_uint64, C.POINTER(Sink), C.POINTER(Camera), C.c_int,
                                       C.POINTER(C.c_float), C.POINTER(BuildStats)]
        h.rtwh_camera_new.argtypes = [C.POINTER(C.c_float)] * 3 + [C.c_float] * 6 + [C.POINTER(Camera)]
        h.rtwh_perlin_new.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        h.rtwh_load_obj.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        h.rtwh_world_create.argtypes = [C.c_char_p, C.c_float, C.c_uint64]
        h.rtwh_world_create.restype = C.c_void_p
        h.rtwh_world_destroy.argtypes = [C.c_void_p]
        h.rtwh_world_destroy.restype = None
        h.rtwh_world_info.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_int, C.POINTER(C.c_float)]
        h.rtwh_world_flatten.argtypes = [C.c_void_p, C.POINTER(Sink), C.POINTER(BuildStats)]
        h.rtwh_open_image.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_size_t]
        h.rtwh_parse_obj.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        h.rtwh_parse_obj.restype = C.c_longlong
        h.rtwh_progress_image_start.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        h.rtwh_progress_pixel.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.c_void_p, C.c_size_t]
        h.rtwh_progress_image_end.argtypes = [C.c_void_p, C.c_size_t]
        h.rtwh_progress_frame_bound.argtypes = [C.c_uint32, C.c_uint32]
        h.rtwh_progress_frame_bound.restype = C.c_size_t
        h.rtwh_progress_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        h.rtwh_progress_frame.restype = C.c_longlong
        h.rtwh_set_asset_dir(ASSET_DIR.encode())
        _host_lib = h
    return _host_lib


def open_image(path: str) -> np.ndarray:
    """ImageTexture::open (image_texture.rs:23-30) of the host front end: the decoded RGB8 texels [h, w, 3]."""
    w, hgt = C.c_uint32(), C.c_uint32()
    hl = host_lib()
    if hl.rtwh_open_image(path.encode(), C.byref(w), C.byref(hgt), None, 0) < 0:
        raise RtwError(hl.rtwh_capi_error().decode())
    out = np.zeros((hgt.value, w.value, 3), np.uint8)
    if hl.rtwh_open_image(path.encode(), C.byref(w), C.byref(hgt), out.ctypes.data, out.size) < 0:
        raise RtwError(hl.rtwh_capi_error().decode())
    return out


def parse_obj(path: str, mode: int = 0, threads: int = 0):
    """OBJ ingest (triangular.rs:170-260 semantics): mode 0 = the parallel parser of the product path, 1 = the
    single-threaded reference parser.  Returns (triangles, checksum over every output array, seconds)."""
    cs, sec = C.c_uint64(), C.c_double()
    n = host_lib().rtwh_parse_obj(path.encode(), mode, threads, C.byref(cs), C.byref(sec))
    if n < 0:
        raise RtwError(host_lib().rtwh_capi_error().decode())
    return int(n), int(cs.value), float(sec.value)


def progress_frame(accum: np.ndarray, spp: int) -> bytes:
    """A frame as the reference's ProgressMessage stream (lib.rs:128-138) in the host receivers' wire format
    (postcard 0.7 + COBS, include/rtw_sink.h): ImageStart, one Pixel per pixel in lib.rs:58 order, ImageEnd."""
    a = np.ascontiguousarray(accum, np.float32)
    hgt, wid = a.shape[0], a.shape[1]
    hl = host_lib()
    buf = np.zeros(hl.rtwh_progress_frame_bound(wid, hgt), np.uint8)
    n = hl.rtwh_progress_frame(a.ctypes.data, wid, hgt, spp, buf.ctypes.data, buf.size)
    if n < 0:
        raise RtwError(hl.rtwh_capi_error().decode())
    return buf[:n].tobytes()


def scene_names():
    buf = C.create_string_buffer(4096)
    host_lib().rtwh_scene_names(buf, len(buf))
    return buf.value.decode().split("\n")


def camera_new(look_from, look_at, up, vfov, aspect, aperture=0.0, focus_dist=10.0, time0=0.0, time1=1.0) -> Camera:
    """Camera::new (camera.rs:25-64), computed by the C++ host front end."""
    cam = Camera()
    rc = host_lib().rtwh_camera_new(_f3(look_from), _f3(look_at), _f3(up), vfov, aspect, aperture, focus_dist, time0, time1,
                                    C.byref(cam))
    if rc < 0:
        raise RtwError(host_lib().rtwh_capi_error().decode())
    return cam


def perlin_new(seed: int):
    g = np.zeros((256, 3), np.float32)
    p = np.zeros((3, 256), np.int32)
    host_lib().rtwh_perlin_new(seed, g.ctypes.data, p[0].ctypes.data, p[1].ctypes.data, p[2].ctypes.data)
    return g, p


def load_obj(path: str):
    """rtwh::load_mesh: (verts[n,9], normals[n,9] or None, uvs[n,6] or None) in file order."""
    n = C.c_uint32(0)
    hn, hu = C.c_int(0), C.c_int(0)
    h = host_lib()
    if h.rtwh_load_obj(path.encode(), C.byref(n), None, None, None, C.byref(hn), C.byref(hu)) < 0:
        raise RtwError(h.rtwh_capi_error().decode())
    v = np.zeros((n.value, 9), np.float32)
    nr = np.zeros((n.value, 9), np.float32) if hn.value else None
    uv = np.zeros((n.value, 6), np.float32) if hu.value else None
    h.rtwh_load_obj(path.encode(), C.byref(n), v.ctypes.data, nr.ctypes.data if nr is not None else None,
                    uv.ctypes.data if uv is not None else None, None, None)
    return v, nr, uv


class World:
    """Scene::generate (scenes.rs:42-60) done once in the host front end: objects, cameras, background.  Flatten it
    into as many backends as needed with Scene.from_world (generation is not repeated)."""

    def __init__(self, name: str, aspect_ratio: float, seed: int = 1):
        h = host_lib()
        self.h = h.rtwh_world_create(name.encode(), aspect_ratio, seed)
        if not self.h:
            raise RtwError(f"world_create({name!r}): {h.rtwh_capi_error().decode()}")
        cams = (Camera * 64)()
        bg = (C.c_float * 3)()
        n = h.rtwh_world_info(self.h, cams, 64, bg)
        self.name = name
        self.cameras = [Camera.from_buffer_copy(cams[i]) for i in range(min(n, 64))]
        self.background = (bg[0], bg[1], bg[2])

    def close(self):
        if self.h:
            host_lib().rtwh_world_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """A scene handle of one backend (rtw_scene* for the CUDA library)."""

    def __init__(self, backend: Backend, device: int = 0, *, _sink: Optional[Sink] = None):
        self.b = backend
        self._sink = _sink
        self.cameras: list = []
        self.background = (0.0, 0.0, 0.0)
        self.build_stats: Optional[BuildStats] = None
        if _sink is not None:
            self.h = C.c_void_p(_sink.scene)
        else:
            self.h = C.c_void_p()
            backend.check(backend.fn("scene_create")(device, C.byref(self.h)), "scene_create")

    # -- life cycle --------------------------------------------------------------------------------
    def close(self):
        if self._sink is not None:
            host_lib().rtwh_sink_close(C.byref(self._sink))
            self._sink = None
            self.h = None
        elif self.h:
            self.b.fn("scene_destroy")(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_name(cls, backend: Backend, name: str, aspect_ratio: float, seed: int = 1, device: int = 0) -> "Scene":
        """Scene::generate (scenes.rs:42-60) in the C++ host front end, flattened into `backend`."""
        h = host_lib()
        sink = Sink()
        if h.rtwh_sink_open(backend.path.encode(), backend.prefix.encode(), device, C.byref(sink)) < 0:
            raise RtwError(h.rtwh_last_error().decode())
        s = cls(backend, device, _sink=sink)
        cams = (Camera * 64)()
        bg = (C.c_float * 3)()
        st = BuildStats()
        n = h.rtwh_build_scene(name.encode(), aspect_ratio, seed, C.byref(sink), cams, 64, bg, C.byref(st))
        if n < 0:
            msg = h.rtwh_capi_error().decode()
            s.close()
            raise RtwError(f"build_scene({name!r}): {msg}")
        s.cameras = [Camera.from_buffer_copy(cams[i]) for i in range(min(n, 64))]
        s.background = (bg[0], bg[1], bg[2])
        s.build_stats = st
        return s

    @classmethod
    def from_world(cls, backend: Backend, world: "World", device: int = 0) -> "Scene":
        """Flatten an already generated World into `backend` and build it (the second half of from_name)."""
        h = host_lib()
        sink = Sink()
        if h.rtwh_sink_open(backend.path.encode(), backend.prefix.encode(), device, C.byref(sink)) < 0:
            raise RtwError(h.rtwh_last_error().decode())
        s = cls(backend, device, _sink=sink)
        st = BuildStats()
        if h.rtwh_world_flatten(world.h, C.byref(sink), C.byref(st)) < 0:
            msg = h.rtwh_capi_error().decode()
            s.close()
            raise RtwError(f"world_flatten({world.name!r}): {msg}")
        s.cameras = list(world.cameras)
        s.background = world.background
        s.build_stats = st
        return s

    def clone(self, device: int) -> "Scene":
        """rtw_scene_clone: a replica of this built scene on another device (buffers copied device to device)."""
        h = C.c_void_p()
        self.b.check(self.b.fn("scene_clone")(self.h, device, C.byref(h)), "scene_clone")
        s = Scene.__new__(Scene)
        s.b, s._sink, s.h = self.b, None, h
        s.cameras, s.background, s.build_stats = list(self.cameras), self.background, self.build_stats
        return s

    # -- emit calls (mirror of rtw_cuda.h) ------------------------------------------------------------
    def _c(self, name, *args):
        return self.b.check(self.b.fn(name)(self.h, *args), name)

    def texture_solid(self, r, g, b): return self._c("add_texture_solid", r, g, b)
    def texture_checker(self, odd, even, freq): return self._c("add_texture_checker", odd, even, freq)
    def texture_uvdebug(self): return self._c("add_texture_uvdebug")

    def texture_noise(self, gradients, perms, scale):
        g = np.ascontiguousarray(gradients, np.float32)
        p = np.ascontiguousarray(perms, np.int32)
        return self._c("add_texture_noise", g.ctypes.data, p[0].ctypes.data, p[1].ctypes.data, p[2].ctypes.data, scale)

    def texture_image(self, rgb):
        a = np.ascontiguousarray(rgb, np.uint8)
        return self._c("add_texture_image", a.ctypes.data, a.shape[1], a.shape[0])

    def lambertian(self, tex): return self._c("add_material_lambertian", tex)
    def lambertian_rgb(self, r, g, b): return self.lambertian(self.texture_solid(r, g, b))
    def metal(self, r, g, b, fuzz): return self._c("add_material_metal", r, g, b, fuzz)
    def dielectric(self, ir): return self._c("add_material_dielectric", ir)
    def diffuse_light(self, tex): return self._c("add_material_diffuse_light", tex)
    def diffuse_light_rgb(self, r, g, b): return self.diffuse_light(self.texture_solid(r, g, b))
    def push_translation(self, off): return self._c("push_translation", _f3(off))
    def push_rotation_y(self, deg): return self._c("push_rotation_y", deg)
    def push_rotation_y_sincos(self, s, c): return self._c("push_rotation_y_sincos", s, c)
    def pop_transform(self): return self._c("pop_transform")
    def begin_group(self): return self._c("begin_group")
    def end_group(self): return self._c("end_group")
    def begin_medium(self, density, tex): return self._c("begin_medium", density, tex)
    def end_medium(self): return self._c("end_medium")
    def sphere(self, c, r, m): return self._c("add_sphere", _f3(c), r, m)
    def moving_sphere(self, c0, t0, c1, t1, r, m): return self._c("add_moving_sphere", _f3(c0), t0, _f3(c1), t1, r, m)
    def xy_rect(self, x0, x1, y0, y1, k, m): return self._c("add_xy_rect", x0, x1, y0, y1, k, m)
    def xz_rect(self, x0, x1, z0, z1, k, m): return self._c("add_xz_rect", x0, x1, z0, z1, k, m)
    def yz_rect(self, y0, y1, z0, z1, k, m): return self._c("add_yz_rect", y0, y1, z0, z1, k, m)
    def cuboid(self, p0, p1, m): return self._c("add_cuboid", _f3(p0), _f3(p1), m)

    def triangles(self, verts, material, normals=None, uvs=None, material_ids=None):
        v = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
        n = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 9)
        t = None if uvs is None else np.ascontiguousarray(uvs, np.float32).reshape(-1, 6)
        mi = None if material_ids is None else np.ascontiguousarray(material_ids, np.int32)
        return self._c("add_triangles", v.shape[0], v.ctypes.data, None if n is None else n.ctypes.data,
                       None if t is None else t.ctypes.data, None if mi is None else mi.ctypes.data, material)

    def build(self, time0=0.0, time1=1.0) -> BuildStats:
        st = BuildStats()
        self._c("build", time0, time1, C.byref(st))
        self.build_stats = st
        return st

    # -- queries -----------------------------------------------------------------------------------------
    @property
    def num_prims(self) -> int:
        return self._c("scene_num_prims")

    def prim_info(self, prim_id):
        t, i, m = C.c_int32(), C.c_int32(), C.c_int32()
        self._c("scene_prim_info", prim_id, C.byref(t), C.byref(i), C.byref(m))
        return t.value, i.value, m.value

    def instance_ops(self, inst):
        kinds = np.zeros(8, np.int32)
        abc = np.zeros((8, 3), np.float32)
        n = self._c("scene_instance_ops", inst, 8, kinds.ctypes.data, abc.ctypes.data)
        return [(int(kinds[k]), tuple(float(x) for x in abc[k])) for k in range(n)]

    def get_bvh(self):
        nn = self._c("scene_num_nodes")
        nodes = np.zeros(2 * nn, BVH_NODE_DTYPE)
        slots = np.zeros(self.num_prims, np.int32)
        root = np.zeros(6, np.float32)
        self._c("get_bvh", nodes.ctypes.data, slots.ctypes.data, root.ctypes.data)
        return nodes, slots, root

    def trace_closest(self, rays: np.ndarray, mode: int = RTW_TRACE_BVH) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        self._c("trace_closest", rays.ctypes.data, rays.shape[0], hits.ctypes.data, mode)
        return hits

    def params(self, width, height, spp, *, max_depth=50, background=None, seed=0, sample_begin=0, sample_end=0,
               tile_size=0, part_rank=0, part_count=0, pool_size=0, slices=0, flags=0, gpus=0) -> RenderParams:
        bg = self.background if background is None else background
        p = RenderParams(width=width, height=height, spp=spp, max_depth=max_depth, sample_begin=sample_begin,
                         sample_end=sample_end, seed=seed, tile_size=tile_size, part_rank=part_rank,
                         part_count=part_count, pool_size=pool_size, slices=slices, flags=flags, gpus=gpus)
        p.background[0], p.background[1], p.background[2] = bg
        return p

    def render(self, cam: Camera, params: RenderParams, out: Optional[np.ndarray] = None):
        """rtw_render with HOST buffers: returns (accum[h, w, 3] float32, RenderStats).  `out`: an existing host frame
        (C-contiguous float32 [h, w, 3]) to render into instead of a new array."""
        accum = np.zeros((params.height, params.width, 3), np.float32) if out is None else out
        assert accum.dtype == np.float32 and accum.flags.c_contiguous and accum.size == params.height * params.width * 3
        st = RenderStats()
        self._c("render", C.byref(cam), C.byref(params), accum.ctypes.data, C.byref(st))
        return accum, st

    def render_device(self, cam: Camera, params: RenderParams, d_accum_ptr: int, stream: int = 0) -> RenderStats:
        """rtw_render_device: d_accum_ptr = device pointer to width*height*3 floats."""
        st = RenderStats()
        self._c("render_device", C.byref(cam), C.byref(params), C.c_void_p(d_accum_ptr), C.c_void_p(stream), C.byref(st))
        return st

    def render_device_stats(self, cam: Camera, params: RenderParams) -> RenderStats:
        """rtw_render_device into a scratch frame allocated for the call (host gets only the stats)."""
        import torch

        buf = torch.empty(params.width * params.height * 3, dtype=torch.float32, device=torch.device("cuda", 0))
        st = self.render_device(cam, params, buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return st

    def render_frames(self, cams, params: RenderParams, on_frame=None) -> int:
        """rtw_render_frames: frame i = cams[i], seed params.seed + i, over the resident scene.
        on_frame(frame_no, accum[h, w, 3] (a copy), stats dict): return False to stop the animation (anything else,
        including None, continues); called in frame order on a helper thread while the next frame renders.
        Returns the number of frames delivered."""
        arr = (Camera * len(cams))(*cams)
        h, w = params.height, params.width
        errors = []

        def tramp(_user, frame, accum, stats):
            try:
                a = np.ctypeslib.as_array(accum, shape=(h, w, 3)).copy()
                return 0 if (on_frame(int(frame), a, stats.contents.as_dict()) is not False) else 1
            except Exception as e:  # never unwind through the C ABI
                errors.append(e)
                return 1

        cb = FRAME_CALLBACK(tramp) if on_frame else C.cast(None, FRAME_CALLBACK)
        n = self._c("render_frames", arr, len(cams), C.byref(params), cb, None)
        if errors:
            raise errors[0]
        return n

    def resolve_rgb8(self, accum: np.ndarray, spp: int) -> np.ndarray:
        a = np.ascontiguousarray(accum, np.float32)
        out = np.zeros(a.shape, np.uint8)
        self._c("resolve_rgb8", a.ctypes.data, a.shape[1], a.shape[0], spp, out.ctypes.data)
        return out


def make_rays(origins, directions, time=0.0, t_min=0.001, t_max=np.inf) -> np.ndarray:
    o = np.asarray(origins, np.float32).reshape(-1, 3)
    d = np.asarray(directions, np.float32).reshape(-1, 3)
    n = max(o.shape[0], d.shape[0])
    r = np.zeros(n, RAY_DTYPE)
    r["origin"] = o
    r["direction"] = d
    r["time"] = time
    r["t_min"] = t_min
    r["t_max"] = t_max
    return r


def read_rtwi(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        assert f.read(4) == b"RTWI"
        ver, w, h = np.frombuffer(f.read(12), "<u4")
        return np.frombuffer(f.read(int(w) * int(h) * 3), np.uint8).reshape(int(h), int(w), 3)


def read_rtwm(path: str):
    with open(path, "rb") as f:
        assert f.read(4) == b"RTWM"
        ver, n, flags = (int(x) for x in np.frombuffer(f.read(12), "<u4"))
        v = np.frombuffer(f.read(n * 36), "<f4").reshape(n, 9)
        nr = np.frombuffer(f.read(n * 36), "<f4").reshape(n, 9) if flags & 1 else None
        uv = np.frombuffer(f.read(n * 24), "<f4").reshape(n, 6) if flags & 2 else None
        return v, nr, uv
