"""ctypes binding of the TEST-ONLY CPU simulator of the device code (tests/hostsim/hostsim.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from nanogi_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhostsim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "libhostsim.so"])
        L = C.CDLL(LIB_PATH)
        L.sim_scene_create.restype = C.c_void_p
        L.sim_scene_create.argtypes = [C.c_void_p]
        L.sim_scene_destroy.argtypes = [C.c_void_p]
        L.sim_scene_destroy.restype = None
        L.sim_last_error.restype = C.c_char_p
        L.sim_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        L.sim_scene_info.restype = None
        L.sim_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        L.sim_trace_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.sim_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sim_eval_bsdf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.sim_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sim_philox.restype = None
        _lib = L
    return _lib


class SimScene:
    def __init__(self, scene_data: capi.SceneData):
        self.L = lib()
        self.scene_data = scene_data
        d = scene_data.desc()
        self.h = self.L.sim_scene_create(C.byref(d))
        if not self.h:
            raise RuntimeError("sim_scene_create: " + self.L.sim_last_error().decode())

    def close(self):
        if self.h:
            self.L.sim_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        out = np.zeros(5)
        self.L.sim_scene_info(self.h, out.ctypes.data)
        return {"n": int(out[0]), "nodes8": int(out[1]), "depth8": int(out[2]), "nodes2": int(out[3]), "pad": out[4]}

    def trace(self, rays, any_hit=False, accel=0):
        rays = np.ascontiguousarray(rays)
        hits = np.empty(rays.shape[0], capi.HIT_DTYPE)
        self.L.sim_trace(self.h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, int(any_hit), accel)
        return hits

    def trace_stats(self, rays, any_hit=False):
        """BVH8 traversal cost of a ray batch: (node steps per ray, triangle tests per ray)."""
        rays = np.ascontiguousarray(rays)
        out = np.zeros(2, np.float64)
        self.L.sim_trace_stats(self.h, rays.ctypes.data, rays.shape[0], int(any_hit), out.ctypes.data)
        return out[0] / max(rays.shape[0], 1), out[1] / max(rays.shape[0], 1)

    def render(self, renderer, num_samples, width, height, max_num_vertices=-1, seed=1, sample_offset=0, film_norm_samples=None,
               wave_capacity=4096, flags=0):
        p = capi.NgiRenderParams()
        p.struct_size = C.sizeof(capi.NgiRenderParams)
        p.renderer = capi.RENDERERS[renderer]
        p.num_samples = num_samples
        p.sample_offset = sample_offset
        p.film_norm_samples = num_samples if film_norm_samples is None else film_norm_samples
        p.max_num_vertices = max_num_vertices
        p.width, p.height = width, height
        p.seed = seed
        p.wave_capacity = wave_capacity
        p.flags = flags
        film = np.zeros((height, width, 3), np.float32)
        stats = np.zeros(8)
        self.L.sim_render(self.h, C.byref(p), film.ctypes.data, stats.ctypes.data)
        return film, {"paths": stats[0], "extend_rays": stats[1], "shadow_rays": stats[2], "iterations": stats[3],
                      # BVH8 traversal cost inside the render (pt / ptdirect / lt / ltdirect): node steps and triangle tests
                      "extend_node_steps": stats[4], "extend_tri_tests": stats[5], "shadow_node_steps": stats[6], "shadow_tri_tests": stats[7]}

    def eval_bsdf(self, queries, wo_in, force_degenerated):
        queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 16)
        n = queries.shape[0]
        wo = np.ascontiguousarray(wo_in if wo_in is not None else np.zeros((n, 3)), dtype=np.float32)
        out = np.empty((n, 8), np.float32)
        self.L.sim_eval_bsdf(self.h, queries.ctypes.data, wo.ctypes.data, n, int(force_degenerated), out.ctypes.data)
        return out


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().sim_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out
