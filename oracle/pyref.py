"""ctypes binding of oracle/_ref/libnanogi_ref.so — the REFERENCE'S OWN CODE (src/nanogi.cpp + include/nanogi/*.hpp, compiled by
oracle/build_ref.sh against the stand-in third-party headers of oracle/refshim/). TEST INFRASTRUCTURE: used by tests/test_reference_pin.py
(and optionally bench.py's CPU arm) to pin the oracle restatement against the real thing. Mirrors the method names of
pyoracle.OracleScene so that the two can be compared call by call."""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libnanogi_ref.so")
BIN_PATH = os.path.join(HERE, "_ref", "nanogi_ref")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `bash oracle/build_ref.sh` in a container that has /root/reference")
        L = C.CDLL(LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_scene_load.argtypes = [C.c_char_p, C.c_double]
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p]
        L.ref_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_visible.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_sample_direction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.ref_sample_direction.restype = None
        L.ref_evaluate_direction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_evaluate_direction.restype = None
        L.ref_sample_position.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.ref_sample_position.restype = None
        L.ref_raster_position.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_raster_position.restype = None
        L.ref_geometry_term.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_geometry_term.restype = C.c_double
        L.ref_orthonormal_basis.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_orthonormal_basis.restype = None
        L.ref_random_stream.argtypes = [C.c_uint, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_random_stream.restype = None
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefScene:
    """The reference's `Scene`, loaded by ITS loader from a schema.yml file (pass a path) or from a scene spec of
    nanogi_b200.scenes (written to a temporary directory as YAML + OBJ first)."""

    def __init__(self, scene, aspect: float = 1.0):
        self.L = lib()
        self._tmp = None
        if not isinstance(scene, str):
            from nanogi_b200 import scenes
            self._tmp = tempfile.TemporaryDirectory(prefix="ngi_ref_scene_")
            scene = scenes.write_scene_files(scene, self._tmp.name)
        self.path = scene
        self.h = self.L.ref_scene_load(scene.encode(), float(aspect))
        if not self.h:
            raise RuntimeError("ref_scene_load: " + self.L.ref_last_error().decode())

    def close(self):
        if self.h:
            self.L.ref_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        out = np.zeros(4)
        self.L.ref_scene_info(self.h, out.ctypes.data)
        return {"prims": int(out[0]), "lights": int(out[1]), "sensor": int(out[2]), "tris": int(out[3])}

    def render(self, renderer, num_samples, width, height, max_num_vertices=-1, seed=1, num_threads=1):
        """Renderer::Render. `seed` is what std::time(nullptr) returns to the reference's master RNG (src/nanogi.cpp:190).
        Returns the film float64 [H, W, 3], row 0 = bottom."""
        r = {"pt": 0, "ptdirect": 1, "lt": 2, "ltdirect": 3, "bdpt": 4}[renderer] if isinstance(renderer, str) else int(renderer)
        film = np.zeros((height, width, 3), np.float64)
        if self.L.ref_render(self.h, r, int(num_samples), int(max_num_vertices), int(width), int(height), int(num_threads), int(seed), film.ctypes.data) != 0:
            raise RuntimeError("ref_render: " + self.L.ref_last_error().decode())
        return film

    def intersect(self, o, d):
        out = np.zeros(19)
        o, d = _d(o), _d(d)
        self.L.ref_intersect(self.h, o.ctypes.data, d.ctypes.data, out.ctypes.data)
        if out[0] == 0:
            return None
        return {"prim": int(out[1]), "p": out[2:5].copy(), "gn": out[5:8].copy(), "sn": out[8:11].copy(), "dpdu": out[11:14].copy(),
                "dpdv": out[14:17].copy(), "uv": out[17:19].copy()}

    def visible(self, p1, p2):
        p1, p2 = _d(p1), _d(p2)
        return bool(self.L.ref_visible(self.h, p1.ctypes.data, p2.ctypes.data))

    def sample_direction(self, prim, query_type, sn, gn, wi, u0, u1, ucomp, p=(0, 0, 0)):
        g = _d(np.concatenate([sn, gn, p])); wi = _d(wi); out = np.zeros(3)
        self.L.ref_sample_direction(self.h, prim, query_type, g.ctypes.data, wi.ctypes.data, u0, u1, ucomp, out.ctypes.data)
        return out.copy()

    def evaluate_direction(self, prim, query_type, sn, gn, wi, wo, trans_dir_el=True, force_degenerated=True, p=(0, 0, 0)):
        g = _d(np.concatenate([sn, gn, p])); wi = _d(wi); wo = _d(wo); out = np.zeros(4)
        self.L.ref_evaluate_direction(self.h, prim, query_type, g.ctypes.data, wi.ctypes.data, wo.ctypes.data, int(trans_dir_el), int(force_degenerated), out.ctypes.data)
        return out[:3].copy(), float(out[3])

    def sample_position(self, prim, u0, u1):
        out = np.zeros(12)
        self.L.ref_sample_position(self.h, prim, u0, u1, out.ctypes.data)
        return {"p": out[0:3].copy(), "gn": out[3:6].copy(), "sn": out[6:9].copy(), "pdf": float(out[9]), "uv": out[10:12].copy()}

    def raster_position(self, prim, wo, w, h):
        out = np.zeros(4); wo = _d(wo)
        self.L.ref_raster_position(self.h, prim, wo.ctypes.data, w, h, out.ctypes.data)
        return bool(out[0]), out[1], out[2], int(out[3])


def geometry_term(p1, sn1, deg1, p2, sn2, deg2):
    a, b, c, d = _d(p1), _d(sn1), _d(p2), _d(sn2)
    return lib().ref_geometry_term(a.ctypes.data, b.ctypes.data, int(deg1), c.ctypes.data, d.ctypes.data, int(deg2))


def orthonormal_basis(a):
    a = _d(a); b = np.zeros(3); c = np.zeros(3)
    lib().ref_orthonormal_basis(a.ctypes.data, b.ctypes.data, c.ctypes.data)
    return b, c


def random_stream(seed, n):
    out = np.zeros(n); nxt = C.c_uint()
    lib().ref_random_stream(int(seed), n, out.ctypes.data, C.byref(nxt))
    return out, nxt.value
