"""ctypes view of the C ABI in include/nanogi_gpu.h and include/nanogi_host.h.

This is binding glue for tests and bench.py only: the product is libnanogi_gpu.so (CUDA, sm_100a) and
the C++ `nanogi` front end. Nothing here computes; there is NO CPU fallback — if libnanogi_gpu.so is
missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB_PATH = os.path.join(_HERE, "libnanogi_gpu.so")
if os.environ.get("NGI_GPU_LIB"):   # A/B experiments with alternative builds of the SAME CUDA module (tools/)
    GPU_LIB_PATH = os.environ["NGI_GPU_LIB"]
HOST_LIB_PATH = os.path.join(_HERE, "libnanogi_host.so")

# ---- enums (include/nanogi_gpu.h) -------------------------------------------------------------
TYPE_D, TYPE_G, TYPE_S, TYPE_L, TYPE_E = 1, 2, 4, 8, 16
TYPE_BSDF_MASK = TYPE_D | TYPE_G | TYPE_S
L_AREA, L_POINT, L_DIRECTIONAL = 0, 1, 2
E_AREA, E_PINHOLE = 0, 1
S_REFLECTION, S_REFRACTION, S_FRESNEL = 0, 1, 2
RENDERER_PT, RENDERER_PTDIRECT, RENDERER_LT, RENDERER_LTDIRECT, RENDERER_BDPT = 0, 1, 2, 3, 4
RENDERERS = {"pt": RENDERER_PT, "ptdirect": RENDERER_PTDIRECT, "lt": RENDERER_LT, "ltdirect": RENDERER_LTDIRECT, "bdpt": RENDERER_BDPT}
RENDER_TIME_KERNELS = 1  # NGI_RENDER_TIME_KERNELS
RENDER_PER_RAY_TRACE = 2  # NGI_RENDER_PER_RAY_TRACE
RENDER_BDPT_PER_THREAD = 4  # NGI_RENDER_BDPT_PER_THREAD
NO_HIT = 0xFFFFFFFF

d3 = C.c_double * 3


class NgiPrimitive(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("first_tri", C.c_int32), ("num_tris", C.c_int32),
        ("l_type", C.c_int32), ("e_type", C.c_int32), ("s_type", C.c_int32),
        ("d_tex", C.c_int32), ("g_tex", C.c_int32),
        ("d_r", d3), ("g_r", d3), ("g_eta", d3), ("g_k", d3), ("g_roughness", C.c_double),
        ("s_r", d3), ("s_eta1", C.c_double), ("s_eta2", C.c_double),
        ("l_le", d3), ("l_vec", d3),
        ("e_position", d3), ("e_vx", d3), ("e_vy", d3), ("e_vz", d3),
        ("e_fov", C.c_double), ("e_aspect", C.c_double), ("e_we", d3),
    ]


class NgiTexture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_float))]


class NgiSceneDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_prims", C.c_uint32), ("num_tris", C.c_uint64),
        ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("texcoords", C.POINTER(C.c_float)),
        ("prims", C.POINTER(NgiPrimitive)),
        ("num_textures", C.c_uint32), ("reserved0", C.c_uint32), ("textures", C.POINTER(NgiTexture)),
    ]


class NgiRenderParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("renderer", C.c_int32),
        ("num_samples", C.c_int64), ("sample_offset", C.c_int64), ("film_norm_samples", C.c_int64),
        ("max_num_vertices", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("accumulate", C.c_int32),
        ("seed", C.c_uint64), ("wave_capacity", C.c_uint32), ("flags", C.c_uint32),
    ]


class NgiRenderStats(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("wave_iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("gpu_seconds", C.c_double), ("trace_kernel_seconds", C.c_double),
        ("logic_kernel_seconds", C.c_double), ("extend_kernel_seconds", C.c_double), ("shadow_kernel_seconds", C.c_double),
        ("logic_launches", C.c_uint64), ("extend_launches", C.c_uint64), ("shadow_launches", C.c_uint64),
        ("reduce_seconds", C.c_double),
    ]


class NgiCommId(C.Structure):
    _fields_ = [("bytes", C.c_char * 128)]


class NgiSceneInfo(C.Structure):
    _fields_ = [
        ("num_tris", C.c_uint64), ("bvh8_nodes", C.c_uint64), ("bvh2_nodes", C.c_uint64), ("device_bytes", C.c_uint64),
        ("build_gpu_seconds", C.c_double), ("scene_min", C.c_float * 3), ("scene_max", C.c_float * 3),
        ("num_lights", C.c_uint32), ("bvh8_max_depth", C.c_uint32), ("bvh2_max_depth", C.c_uint32), ("reserved0", C.c_uint32),
    ]


class NgiCliOptions(C.Structure):
    _fields_ = [
        ("help", C.c_int32), ("has_scene", C.c_int32), ("has_renderer", C.c_int32), ("has_num_threads", C.c_int32),
        ("has_seed", C.c_int32),
        ("scene", C.c_char * 1024), ("result", C.c_char * 1024), ("renderer", C.c_char * 64),
        ("num_samples", C.c_int64), ("max_num_vertices", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("num_threads", C.c_int32), ("grain_size", C.c_int64), ("progress_update_interval", C.c_int64),
        ("render_time", C.c_double), ("progress_image_update_interval", C.c_double),
        ("progress_image_update_format", C.c_char * 1024),
        ("gpus", C.c_int32), ("wave_capacity", C.c_uint32), ("seed", C.c_uint64), ("device", C.c_char * 16),
        ("sample_offset", C.c_int64), ("resume_from", C.c_char * 1024),
    ]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("tri", np.uint32)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16

# Every symbol include/nanogi_gpu.h declares (checked by tests/test_cabi_symbols.py)
GPU_SYMBOLS = [
    "ngi_gpu_device_count", "ngi_gpu_scene_create", "ngi_gpu_scene_info", "ngi_gpu_scene_destroy",
    "ngi_gpu_render", "ngi_gpu_render_device", "ngi_gpu_trace", "ngi_gpu_trace_device",
    "ngi_gpu_eval_bsdf", "ngi_gpu_last_error", "ngi_gpu_abi_version",
    "ngi_gpu_shard_range", "ngi_gpu_group_create", "ngi_gpu_group_render", "ngi_gpu_group_scene", "ngi_gpu_group_destroy",
    "ngi_gpu_comm_get_id", "ngi_gpu_comm_create", "ngi_gpu_comm_reduce_film", "ngi_gpu_comm_destroy",
]
ABI_VERSION = 2
ERR_NCCL = -6
HOST_SYMBOLS = [
    "ngi_host_scene_load", "ngi_host_scene_desc", "ngi_host_scene_sensor", "ngi_host_scene_num_lights",
    "ngi_host_scene_free", "ngi_host_save_image", "ngi_host_load_image", "ngi_host_parse_cli", "ngi_host_usage", "ngi_host_last_error",
]


# ---- scene container --------------------------------------------------------------------------
@dataclass
class SceneData:
    """Numpy-side owner of the arrays a NgiSceneDesc points into."""
    positions: np.ndarray                     # float32 [nTri, 3, 3]
    normals: np.ndarray                       # float32 [nTri, 3, 3]
    prims: List[NgiPrimitive]
    texcoords: Optional[np.ndarray] = None    # float32 [nTri, 3, 2]
    name: str = "scene"
    textures: List[np.ndarray] = field(default_factory=list)   # float32 [H, W, 3] each, row 0 = top (rt.hpp:157-270)
    _keep: list = field(default_factory=list, repr=False)

    @property
    def num_tris(self) -> int:
        return int(self.positions.shape[0])

    def desc(self) -> NgiSceneDesc:
        self.positions = np.ascontiguousarray(self.positions, dtype=np.float32).reshape(-1, 3, 3)
        self.normals = np.ascontiguousarray(self.normals, dtype=np.float32).reshape(-1, 3, 3)
        assert self.positions.shape == self.normals.shape
        arr = (NgiPrimitive * len(self.prims))(*self.prims)
        d = NgiSceneDesc()
        d.struct_size = C.sizeof(NgiSceneDesc)
        d.num_prims = len(self.prims)
        d.num_tris = self.num_tris
        d.positions = self.positions.ctypes.data_as(C.POINTER(C.c_float))
        d.normals = self.normals.ctypes.data_as(C.POINTER(C.c_float))
        if self.texcoords is not None:
            self.texcoords = np.ascontiguousarray(self.texcoords, dtype=np.float32).reshape(-1, 3, 2)
            d.texcoords = self.texcoords.ctypes.data_as(C.POINTER(C.c_float))
        d.prims = arr
        self.textures = [np.ascontiguousarray(t, dtype=np.float32) for t in self.textures]
        tex = (NgiTexture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            tex[i].height, tex[i].width = int(t.shape[0]), int(t.shape[1])
            tex[i].rgb = t.ctypes.data_as(C.POINTER(C.c_float))
        d.num_textures = len(self.textures)
        d.textures = tex if self.textures else None
        self._keep = [arr, tex]
        return d

    def sensor_prim(self) -> int:
        idx = -1
        for i, p in enumerate(self.prims):
            if p.type & TYPE_E:
                idx = i
        return idx

    def light_prims(self) -> List[int]:
        return [i for i, p in enumerate(self.prims) if p.type & TYPE_L]

    def set_aspect(self, aspect: float) -> None:
        for p in self.prims:
            if p.type & TYPE_E:
                p.e_aspect = aspect


def make_prim(**kw) -> NgiPrimitive:
    p = NgiPrimitive()
    p.first_tri = -1
    p.d_tex = -1
    p.g_tex = -1
    for k, v in kw.items():
        if isinstance(getattr(p, k), C.Array):
            setattr(p, k, d3(*[float(x) for x in v]))
        else:
            setattr(p, k, v)
    return p


# ---- libraries --------------------------------------------------------------------------------
_host = None
_gpu = None


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(HOST_LIB_PATH)
        lib.ngi_host_scene_load.argtypes = [C.c_char_p, C.c_double, C.POINTER(C.c_void_p)]
        lib.ngi_host_scene_desc.argtypes = [C.c_void_p, C.POINTER(NgiSceneDesc)]
        lib.ngi_host_scene_sensor.argtypes = [C.c_void_p]
        lib.ngi_host_scene_num_lights.argtypes = [C.c_void_p]
        lib.ngi_host_scene_free.argtypes = [C.c_void_p]
        lib.ngi_host_scene_free.restype = None
        lib.ngi_host_save_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        lib.ngi_host_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_uint64]
        lib.ngi_host_parse_cli.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(NgiCliOptions)]
        lib.ngi_host_usage.restype = C.c_char_p
        lib.ngi_host_last_error.restype = C.c_char_p
        _host = lib
    return _host


def gpu_lib():
    """Loads libnanogi_gpu.so. Raises when it is missing — the product path has no fallback."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(GPU_LIB_PATH):
            raise RuntimeError(f"{GPU_LIB_PATH} missing: the CUDA module is required (no CPU fallback); "
                               "build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(GPU_LIB_PATH)
        lib.ngi_gpu_scene_create.argtypes = [C.POINTER(NgiSceneDesc), C.c_int, C.POINTER(C.c_void_p)]
        lib.ngi_gpu_scene_info.argtypes = [C.c_void_p, C.POINTER(NgiSceneInfo)]
        lib.ngi_gpu_scene_destroy.argtypes = [C.c_void_p]
        lib.ngi_gpu_scene_destroy.restype = None
        lib.ngi_gpu_render.argtypes = [C.c_void_p, C.POINTER(NgiRenderParams), C.c_void_p, C.POINTER(NgiRenderStats)]
        lib.ngi_gpu_render_device.argtypes = [C.c_void_p, C.POINTER(NgiRenderParams), C.c_void_p, C.c_void_p, C.POINTER(NgiRenderStats)]
        lib.ngi_gpu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        lib.ngi_gpu_trace_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        lib.ngi_gpu_eval_bsdf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        lib.ngi_gpu_last_error.restype = C.c_char_p
        lib.ngi_gpu_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.ngi_gpu_shard_range.restype = None
        lib.ngi_gpu_group_create.argtypes = [C.POINTER(NgiSceneDesc), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        lib.ngi_gpu_group_render.argtypes = [C.c_void_p, C.POINTER(NgiRenderParams), C.c_void_p, C.POINTER(NgiRenderStats)]
        lib.ngi_gpu_group_scene.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.ngi_gpu_group_destroy.argtypes = [C.c_void_p]
        lib.ngi_gpu_group_destroy.restype = None
        lib.ngi_gpu_comm_get_id.argtypes = [C.POINTER(NgiCommId)]
        lib.ngi_gpu_comm_create.argtypes = [C.POINTER(NgiCommId), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.ngi_gpu_comm_reduce_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        lib.ngi_gpu_comm_destroy.argtypes = [C.c_void_p]
        lib.ngi_gpu_comm_destroy.restype = None
        _gpu = lib
    return _gpu


class NgiError(RuntimeError):
    pass


def _check(status: int, what: str):
    if status != 0:
        raise NgiError(f"{what} failed ({status}): {gpu_lib().ngi_gpu_last_error().decode()}")


# ---- host front end ---------------------------------------------------------------------------
def load_scene_file(path: str, aspect: float) -> SceneData:
    """Scene::Load through the C++ front end; the arrays are copied out into numpy."""
    lib = host_lib()
    h = C.c_void_p()
    if lib.ngi_host_scene_load(path.encode(), float(aspect), C.byref(h)) != 0:
        raise NgiError(f"scene load failed: {lib.ngi_host_last_error().decode()}")
    try:
        d = NgiSceneDesc()
        lib.ngi_host_scene_desc(h, C.byref(d))
        n = int(d.num_tris)
        pos = np.ctypeslib.as_array(d.positions, shape=(n, 3, 3)).copy() if n else np.zeros((0, 3, 3), np.float32)
        nrm = np.ctypeslib.as_array(d.normals, shape=(n, 3, 3)).copy() if n else np.zeros((0, 3, 3), np.float32)
        uv = np.ctypeslib.as_array(d.texcoords, shape=(n, 3, 2)).copy() if (n and d.texcoords) else None
        prims = []
        for i in range(d.num_prims):
            p = NgiPrimitive()
            C.memmove(C.byref(p), C.byref(d.prims[i]), C.sizeof(NgiPrimitive))
            prims.append(p)
        textures = []
        for i in range(d.num_textures):
            t = d.textures[i]
            textures.append(np.ctypeslib.as_array(t.rgb, shape=(t.height, t.width, 3)).copy())
        return SceneData(pos, nrm, prims, uv, name=os.path.basename(path), textures=textures)
    finally:
        lib.ngi_host_scene_free(h)


def save_image(path: str, film: np.ndarray) -> None:
    film = np.ascontiguousarray(film, dtype=np.float32)
    h, w = film.shape[:2]
    if host_lib().ngi_host_save_image(path.encode(), film.ctypes.data, w, h) != 0:
        raise NgiError(host_lib().ngi_host_last_error().decode())


def load_image(path: str) -> np.ndarray:
    """Texture::Load through the C++ front end: float32 [H, W, 3], row 0 = top."""
    lib = host_lib()
    w, h = C.c_int(), C.c_int()
    if lib.ngi_host_load_image(path.encode(), C.byref(w), C.byref(h), None, 0) != 0:
        raise NgiError(lib.ngi_host_last_error().decode())
    out = np.empty((h.value, w.value, 3), np.float32)
    if lib.ngi_host_load_image(path.encode(), C.byref(w), C.byref(h), out.ctypes.data, out.size) != 0:
        raise NgiError(lib.ngi_host_last_error().decode())
    return out


def parse_cli(argv: List[str]) -> NgiCliOptions:
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    out = NgiCliOptions()
    if host_lib().ngi_host_parse_cli(len(argv), arr, C.byref(out)) != 0:
        raise NgiError(host_lib().ngi_host_last_error().decode())
    return out


# ---- GPU module -------------------------------------------------------------------------------
class GpuScene:
    """RAII wrapper of an ngi_gpu_scene handle."""

    def __init__(self, scene: SceneData, device: int = 0):
        self.lib = gpu_lib()
        self.handle = C.c_void_p()
        self.scene = scene
        d = scene.desc()
        _check(self.lib.ngi_gpu_scene_create(C.byref(d), device, C.byref(self.handle)), "ngi_gpu_scene_create")

    def close(self):
        if self.handle:
            self.lib.ngi_gpu_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> NgiSceneInfo:
        out = NgiSceneInfo()
        _check(self.lib.ngi_gpu_scene_info(self.handle, C.byref(out)), "ngi_gpu_scene_info")
        return out

    @staticmethod
    def _params(renderer, num_samples, width, height, max_num_vertices=-1, seed=1, sample_offset=0,
                film_norm_samples=None, wave_capacity=0, accumulate=0, flags=0) -> NgiRenderParams:
        p = NgiRenderParams()
        p.struct_size = C.sizeof(NgiRenderParams)
        p.renderer = RENDERERS[renderer] if isinstance(renderer, str) else int(renderer)
        p.num_samples = int(num_samples)
        p.sample_offset = int(sample_offset)
        p.film_norm_samples = int(num_samples if film_norm_samples is None else film_norm_samples)
        p.max_num_vertices = int(max_num_vertices)
        p.width, p.height = int(width), int(height)
        p.accumulate = int(accumulate)
        p.seed = int(seed)
        p.wave_capacity = int(wave_capacity)
        p.flags = int(flags)
        return p

    def render(self, renderer, num_samples, width, height, **kw):
        """ngi_gpu_render: film returned as float32 [H, W, 3], row 0 = bottom."""
        p = self._params(renderer, num_samples, width, height, **kw)
        film = np.empty((height, width, 3), np.float32)
        stats = NgiRenderStats()
        _check(self.lib.ngi_gpu_render(self.handle, C.byref(p), film.ctypes.data, C.byref(stats)), "ngi_gpu_render")
        return film, stats

    def render_device(self, film_ptr: int, stream_ptr: int, renderer, num_samples, width, height, **kw):
        p = self._params(renderer, num_samples, width, height, **kw)
        stats = NgiRenderStats()
        _check(self.lib.ngi_gpu_render_device(self.handle, C.byref(p), C.c_void_p(film_ptr), C.c_void_p(stream_ptr), C.byref(stats)),
               "ngi_gpu_render_device")
        return stats

    def trace(self, rays: np.ndarray, any_hit: bool = False, accel: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        _check(self.lib.ngi_gpu_trace(self.handle, rays.ctypes.data, rays.shape[0], hits.ctypes.data, int(any_hit), accel), "ngi_gpu_trace")
        return hits

    def trace_device(self, rays_ptr: int, n: int, hits_ptr: int, any_hit: bool = False, accel: int = 0) -> float:
        sec = C.c_double()
        _check(self.lib.ngi_gpu_trace_device(self.handle, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr), int(any_hit), accel, C.byref(sec)),
               "ngi_gpu_trace_device")
        return sec.value

    def eval_bsdf(self, queries: np.ndarray, wo_in: Optional[np.ndarray], force_degenerated: bool) -> np.ndarray:
        queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 16)
        n = queries.shape[0]
        wo = np.ascontiguousarray(wo_in if wo_in is not None else np.zeros((n, 3)), dtype=np.float32)
        out = np.empty((n, 8), np.float32)
        _check(self.lib.ngi_gpu_eval_bsdf(self.handle, queries.ctypes.data, wo.ctypes.data, n, int(force_degenerated), out.ctypes.data),
               "ngi_gpu_eval_bsdf")
        return out


def _map_torchs_nccl_first():
    """libnanogi_gpu.so opens "libnccl.so.2" with dlopen on first use. In a Python process that ALSO imports torch, torch's own
    (newer) NCCL must be the copy mapped under that soname: if the system library got there first, a later `import torch` fails
    with an undefined NCCL symbol. So the glue imports torch (when installed) before the first NCCL entry point is called."""
    try:
        import torch  # noqa: F401
    except ImportError:
        pass


class GpuGroup:
    """ngi_gpu_group_*: every device of ONE process — scene built once and broadcast, samples sharded by index, one NCCL
    film reduce per render (what `nanogi --gpus N` runs)."""

    def __init__(self, scene: SceneData, devices):
        _map_torchs_nccl_first()
        self.lib = gpu_lib()
        self.handle = C.c_void_p()
        self.scene = scene
        self.devices = list(devices)
        d = scene.desc()
        arr = (C.c_int * len(self.devices))(*self.devices)
        _check(self.lib.ngi_gpu_group_create(C.byref(d), arr, len(self.devices), C.byref(self.handle)), "ngi_gpu_group_create")

    def close(self):
        if self.handle:
            self.lib.ngi_gpu_group_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, renderer, num_samples, width, height, **kw):
        p = GpuScene._params(renderer, num_samples, width, height, **kw)
        film = np.empty((height, width, 3), np.float32)
        stats = NgiRenderStats()
        _check(self.lib.ngi_gpu_group_render(self.handle, C.byref(p), film.ctypes.data, C.byref(stats)), "ngi_gpu_group_render")
        return film, stats


class GpuComm:
    """ngi_gpu_comm_*: this process's rank of an NCCL communicator that spans one process per GPU. `exchange` hands rank 0's
    128-byte id to the other ranks (any host channel: bench.py passes a torch.distributed broadcast)."""

    def __init__(self, rank: int, world_size: int, device: int, exchange):
        _map_torchs_nccl_first()
        self.lib = gpu_lib()
        self.handle = C.c_void_p()
        cid = NgiCommId()
        if rank == 0:
            _check(self.lib.ngi_gpu_comm_get_id(C.byref(cid)), "ngi_gpu_comm_get_id")
        blob = exchange(C.string_at(C.byref(cid), 128) if rank == 0 else None)
        assert len(blob) == 128
        C.memmove(C.byref(cid), blob, 128)
        _check(self.lib.ngi_gpu_comm_create(C.byref(cid), rank, world_size, device, C.byref(self.handle)), "ngi_gpu_comm_create")

    def reduce_film(self, film_ptr: int, num_floats: int, root: int = 0, stream_ptr: int = 0):
        _check(self.lib.ngi_gpu_comm_reduce_film(self.handle, C.c_void_p(film_ptr), num_floats, root, C.c_void_p(stream_ptr)),
               "ngi_gpu_comm_reduce_film")

    def close(self):
        if self.handle:
            self.lib.ngi_gpu_comm_destroy(self.handle)
            self.handle = C.c_void_p()


def shard_range(num_samples: int, rank: int, world_size: int):
    """ngi_gpu_shard_range: (offset, count) of `rank` — the product's own arithmetic (no device needed)."""
    off, cnt = C.c_int64(), C.c_int64()
    gpu_lib().ngi_gpu_shard_range(int(num_samples), int(rank), int(world_size), C.byref(off), C.byref(cnt))
    return off.value, cnt.value


def device_count() -> int:
    return int(gpu_lib().ngi_gpu_device_count())
