"""ctypes binding of the CPU oracle (oracle/oracle.cpp -> oracle/liboracle.so).

TEST INFRASTRUCTURE. Import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg. The product never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.c_void_p]
        L.oracle_scene_destroy.argtypes = [C.c_void_p]
        L.oracle_scene_destroy.restype = None
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        L.oracle_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_visible.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_sample_direction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.oracle_sample_direction.restype = None
        L.oracle_evaluate_direction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_evaluate_direction.restype = None
        L.oracle_sample_position.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.oracle_sample_position.restype = None
        L.oracle_raster_position.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_raster_position.restype = None
        L.oracle_fresnel.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.oracle_fresnel.restype = C.c_double
        L.oracle_geometry_term.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_geometry_term.restype = C.c_double
        L.oracle_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_philox.restype = None
        L.oracle_orthonormal_basis.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_orthonormal_basis.restype = None
        L.oracle_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_scene_info.restype = None
        L.oracle_light_cdf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleScene:
    def __init__(self, scene_data):
        self.L = lib()
        self.scene_data = scene_data
        d = scene_data.desc()
        self.h = self.L.oracle_scene_create(C.byref(d))
        if not self.h:
            raise RuntimeError("oracle_scene_create: " + self.L.oracle_last_error().decode())

    def close(self):
        if self.h:
            self.L.oracle_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, renderer, num_samples, width, height, max_num_vertices=-1, seed=1, rng_mode=0, num_threads=0,
               sample_offset=0, film_norm_samples=None):
        """Returns (film float64 [H, W, 3] row 0 = bottom, stats dict)."""
        r = {"pt": 0, "ptdirect": 1, "lt": 2, "ltdirect": 3, "bdpt": 4}[renderer] if isinstance(renderer, str) else int(renderer)
        film = np.zeros((height, width, 3), np.float64)
        stats = np.zeros(4, np.float64)
        norm = num_samples if film_norm_samples is None else film_norm_samples
        rc = self.L.oracle_render(self.h, r, int(num_samples), int(sample_offset), int(norm), int(max_num_vertices), int(width), int(height),
                                  int(seed), int(rng_mode), int(num_threads), film.ctypes.data, stats.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_render: " + self.L.oracle_last_error().decode())
        return film, {"paths": stats[0], "extend_rays": stats[1], "shadow_rays": stats[2], "seconds": stats[3]}

    def trace(self, rays, mode=0, num_threads=0):
        """mode 0 closest (BVH), 1 any-hit, 2 closest brute force."""
        from nanogi_b200.capi import HIT_DTYPE, RAY_DTYPE
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        rc = self.L.oracle_trace(self.h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, mode, num_threads)
        assert rc == 0
        return hits

    def intersect(self, o, d):
        out = np.zeros(19, np.float64)
        o, d = _d(o), _d(d)
        self.L.oracle_intersect(self.h, o.ctypes.data, d.ctypes.data, out.ctypes.data)
        if out[0] == 0:
            return None
        return {"tri": int(out[1]), "p": out[2:5], "gn": out[5:8], "sn": out[8:11], "dpdu": out[11:14], "dpdv": out[14:17], "uv": out[17:19]}

    def visible(self, p1, p2):
        p1, p2 = _d(p1), _d(p2)
        return bool(self.L.oracle_visible(self.h, p1.ctypes.data, p2.ctypes.data))

    def sample_direction(self, prim, query_type, sn, gn, wi, u0, u1, ucomp, p=(0, 0, 0)):
        g = _d(np.concatenate([sn, gn, p]))
        wi = _d(wi)
        out = np.zeros(4, np.float64)
        self.L.oracle_sample_direction(self.h, prim, query_type, g.ctypes.data, wi.ctypes.data, u0, u1, ucomp, out.ctypes.data)
        return out[:3].copy(), bool(out[3])

    def evaluate_direction(self, prim, query_type, sn, gn, wi, wo, trans_dir_el=True, force_degenerated=True, p=(0, 0, 0)):
        g = _d(np.concatenate([sn, gn, p]))
        wi, wo = _d(wi), _d(wo)
        out = np.zeros(4, np.float64)
        self.L.oracle_evaluate_direction(self.h, prim, query_type, g.ctypes.data, wi.ctypes.data, wo.ctypes.data, int(trans_dir_el),
                                         int(force_degenerated), out.ctypes.data)
        return out[:3].copy(), float(out[3])

    def sample_position(self, prim, u0, u1):
        out = np.zeros(10, np.float64)
        self.L.oracle_sample_position(self.h, prim, u0, u1, out.ctypes.data)
        return {"p": out[0:3].copy(), "gn": out[3:6].copy(), "sn": out[6:9].copy(), "pdf": float(out[9])}

    def raster_position(self, prim, wo, w, h):
        out = np.zeros(4, np.float64)
        wo = _d(wo)
        self.L.oracle_raster_position(self.h, prim, wo.ctypes.data, w, h, out.ctypes.data)
        return bool(out[0]), out[1], out[2], int(out[3])

    def fresnel(self, prim, cos_i, eta_i, eta_t):
        return self.L.oracle_fresnel(self.h, prim, cos_i, eta_i, eta_t)

    def info(self):
        out = np.zeros(6, np.float64)
        self.L.oracle_scene_info(self.h, out.ctypes.data)
        return {"tris": int(out[0]), "prims": int(out[1]), "lights": int(out[2]), "sensor": int(out[3]), "bvh_nodes": int(out[4]), "pad": out[5]}

    def light_cdf(self, prim):
        cdf = np.zeros(1 << 16, np.float64)
        inv = C.c_double()
        n = self.L.oracle_light_cdf(self.h, prim, cdf.ctypes.data, cdf.shape[0], C.byref(inv))
        return cdf[:n].copy(), inv.value


def geometry_term(p1, sn1, deg1, p2, sn2, deg2):
    a, b, c, d = _d(p1), _d(sn1), _d(p2), _d(sn2)
    return lib().oracle_geometry_term(a.ctypes.data, b.ctypes.data, int(deg1), c.ctypes.data, d.ctypes.data, int(deg2))


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def orthonormal_basis(a):
    a = _d(a)
    b, c = np.zeros(3), np.zeros(3)
    lib().oracle_orthonormal_basis(a.ctypes.data, b.ctypes.data, c.ctypes.data)
    return b, c
