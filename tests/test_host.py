"""Host front end (C++): YAML subset, OBJ loader, Scene::Load mirror, CLI parser, film writers."""
import math
import os
import struct
import zlib

import numpy as np
import pytest

from nanogi_b200 import capi, scenes
from tests.conftest import REFERENCE

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")


def test_cli_defaults_match_reference_table():
    o = capi.parse_cli(["nanogi", "pt", "scene.yml"])
    # reference src/nanogi.cpp:2003-2019
    assert o.renderer == b"pt" and o.scene == b"scene.yml" and o.result == b"render.hdr"
    assert o.num_samples == 10000000 and o.max_num_vertices == -1 and o.width == 1280 and o.height == 720
    assert o.grain_size == 10000 and o.progress_update_interval == 100000 and o.render_time == -1
    assert o.progress_image_update_interval == -1 and o.progress_image_update_format == b"progress/{{count}}.png"
    assert o.has_num_threads == 0 and o.gpus == 1


def test_cli_positional_and_short_options():
    o = capi.parse_cli(["nanogi", "ptdirect", "s.yml", "out.exr", "1920", "1080", "-n", "2123366400", "-m", "8", "-j", "-1"])
    assert (o.renderer, o.scene, o.result, o.width, o.height) == (b"ptdirect", b"s.yml", b"out.exr", 1920, 1080)
    assert o.num_samples == 2123366400 and o.max_num_vertices == 8 and o.num_threads == -1 and o.has_num_threads
    # -h is --height (help is --help only), long options with '='
    o = capi.parse_cli(["nanogi", "-r", "pt", "-i", "a.yml", "-o", "b.png", "-w", "64", "-h", "32", "--num-samples=5", "--gpus", "8", "--seed", "42"])
    assert (o.width, o.height, o.num_samples, o.gpus, o.seed, o.has_seed) == (64, 32, 5, 8, 42, 1)
    assert capi.parse_cli(["nanogi"]).help == 1 and capi.parse_cli(["nanogi", "--help"]).help == 1
    with pytest.raises(capi.NgiError, match="renderer"):
        capi.parse_cli(["nanogi", "-n", "10"])
    with pytest.raises(capi.NgiError, match="unrecognised"):
        capi.parse_cli(["nanogi", "pt", "--bogus", "1"])
    with pytest.raises(capi.NgiError, match="invalid"):
        capi.parse_cli(["nanogi", "pt", "s.yml", "o.hdr", "abc"])


def test_scene_files_round_trip(tmp_path):
    spec = scenes.cornell_spheres()
    path = scenes.write_scene_files(spec, str(tmp_path))
    sd_file = capi.load_scene_file(path, 1.5)
    sd_mem = scenes.to_scene_data(spec, 1.5)
    assert sd_file.num_tris == sd_mem.num_tris == 38 + 2 * 1280
    assert np.array_equal(sd_file.positions, sd_mem.positions)
    assert np.allclose(sd_file.normals, sd_mem.normals, atol=1e-6)
    assert len(sd_file.prims) == len(sd_mem.prims)
    for a, b in zip(sd_file.prims, sd_mem.prims):
        for f, _ in capi.NgiPrimitive._fields_:
            va, vb = getattr(a, f), getattr(b, f)
            if hasattr(va, "__len__"):
                assert np.allclose(list(va), list(vb), rtol=1e-12, atol=1e-12), f
            else:
                assert va == pytest.approx(vb, rel=1e-12, abs=1e-12), f
    assert sd_file.sensor_prim() == len(spec) - 1 and sd_file.light_prims() == [0]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference fixtures only exist in the build container")
def test_loads_reference_fixtures():
    sd = capi.load_scene_file(os.path.join(REFERENCE, "utils/runc/data/cornelbox/scene.yml"), 1.0)
    assert sd.num_tris == 38 and len(sd.prims) == 9                      # 38 triangles after triangulation
    assert [p.num_tris for p in sd.prims[:8]] == [2, 2, 4, 2, 2, 2, 12, 12]
    assert sd.prims[0].type == capi.TYPE_L | capi.TYPE_D and list(sd.prims[0].l_le) == [10, 10, 10]
    cam = sd.prims[8]
    assert cam.type == capi.TYPE_E and cam.e_type == capi.E_PINHOLE and list(cam.e_position) == [278, 273, -800]
    assert abs(cam.e_fov - math.radians(39.3077)) < 1e-12 and list(cam.e_vz) == [0, 0, -1]
    # our generated Cornell box is the same scene (same published data; the fixture carries float noise)
    gen = scenes.to_scene_data(scenes.cornell_box(), 1.0)
    assert np.allclose(np.sort(sd.positions.reshape(-1, 3), axis=0), np.sort(gen.positions.reshape(-1, 3), axis=0), atol=2e-3)
    sd2 = capi.load_scene_file(os.path.join(REFERENCE, "utils/runc/data/scene.yml"), 16 / 9)
    assert sd2.num_tris == 3 * 1280 + 3 * 2
    assert [p.type for p in sd2.prims] == [capi.TYPE_L | capi.TYPE_D, capi.TYPE_G, capi.TYPE_D, capi.TYPE_S, capi.TYPE_D, capi.TYPE_D, capi.TYPE_E]
    assert sd2.prims[1].g_roughness == 0.1 and sd2.prims[3].s_type == capi.S_FRESNEL and sd2.prims[3].s_eta2 == 2
    n = sd2.normals[sd2.prims[1].first_tri:sd2.prims[1].first_tri + 1280].reshape(-1, 3)
    assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-3)


def test_yaml_and_loader_errors(tmp_path):
    def write(text, name="scene.yml"):
        p = tmp_path / name
        p.write_text(text)
        return str(p)
    with pytest.raises(capi.NgiError, match="version"):
        capi.load_scene_file(write("version: 9\nscene:\n  primitives: []\n"), 1.0)
    with pytest.raises(capi.NgiError, match="Invalid primitive type"):
        capi.load_scene_file(write("version: 5\nscene:\n  primitives:\n    - type: [L, E]\n      params: {}\n"), 1.0)
    with pytest.raises(capi.NgiError):
        capi.load_scene_file(str(tmp_path / "missing.yml"), 1.0)
    # mesh without normals and without postprocess: clean error instead of the reference's null dereference (rt.hpp:1688)
    (tmp_path / "t.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    y = "version: 5\nscene:\n  primitives:\n    - type: [D]\n      mesh:\n        path: t.obj\n      params:\n        D:\n          R: [1, 1, 1]\n"
    with pytest.raises(capi.NgiError, match="no normals"):
        capi.load_scene_file(write(y), 1.0)
    y2 = y.replace("        path: t.obj\n", "        path: t.obj\n        postprocess:\n          generate_normals: true\n          generate_smooth_normals: false\n")
    y2 += "    - type: [E]\n      params:\n        E:\n          type: pinhole\n          pinhole:\n            We: [1,1,1]\n            view: {eye: [0,0,3], center: [0,0,0], up: [0,1,0]}\n            perspective:\n              fov: 45\n"
    sd = capi.load_scene_file(write(y2), 2.0)
    assert sd.num_tris == 1 and np.allclose(sd.normals[0], [[0, 0, 1]] * 3) and sd.prims[1].e_aspect == 2.0


def test_area_sensor_needs_uv_on_its_own_mesh(tmp_path):
    """rt.hpp:1919-1924 checks the SENSOR's mesh: an earlier mesh with texture coordinates must not satisfy it"""
    (tmp_path / "uv.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1\n")
    (tmp_path / "nouv.obj").write_text("v 0 0 2\nv 1 0 2\nv 0 1 2\nvn 0 0 -1\nf 1//1 2//1 3//1\n")
    y = ("version: 5\nscene:\n  primitives:\n    - type: [D]\n      mesh:\n        path: uv.obj\n      params:\n        D:\n          R: [1, 1, 1]\n"
         "    - type: [E]\n      mesh:\n        path: nouv.obj\n      params:\n        E:\n          type: area\n          area:\n            We: [1, 1, 1]\n")
    p = tmp_path / "scene.yml"
    p.write_text(y)
    with pytest.raises(capi.NgiError, match="UV coordinates"):
        capi.load_scene_file(str(p), 1.0)
    p.write_text(y.replace("nouv.obj", "uv.obj"))
    assert capi.load_scene_file(str(p), 1.0).num_tris == 2


def test_malformed_png_is_rejected(tmp_path):
    import struct, zlib
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    bad = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", b"\0\0\0\1\0\0\0\1") + chunk(b"IEND", b"")       # IHDR of 8 bytes instead of 13
    f = tmp_path / "bad.png"
    f.write_bytes(bad)
    with pytest.raises(capi.NgiError, match="IHDR"):
        capi.load_image(str(f))


def test_obj_features(tmp_path):
    (tmp_path / "m.obj").write_text(
        "# quad + negative indices + vt\no first\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
        "f 1/1/1 2/2/1 3/3/1 4/4/1\nf -4/-4/-1 -3/-3/-1 -2/-2/-1\no second\nv 5 5 5\nv 6 5 5\nv 5 6 5\nf 5//1 6//1 7//1\n")
    y = ("version: 4\nscene:\n  primitives:\n    - type: [D]\n      mesh:\n        path: 'm.obj'\n      params:\n        D:\n          R: [0.5, 0.5, 0.5]\n"
         "    - type: [E]\n      params:\n        E:\n          type: pinhole\n          pinhole:\n            We: [1, 1, 1]\n            view:\n"
         "              eye: [0, 0, 3]\n              center: [0, 0, 0]\n              up: [0, 1, 0]\n            perspective:\n              fov: 45\n")
    (tmp_path / "scene.yml").write_text(y)
    sd = capi.load_scene_file(str(tmp_path / "scene.yml"), 1.0)
    assert sd.num_tris == 3                                # quad -> fan (0,1,2),(0,2,3) + one triangle; 2nd object ignored (mMeshes[0])
    assert np.array_equal(sd.positions[1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    assert sd.texcoords is not None and np.array_equal(sd.texcoords[0], [[0, 0], [1, 0], [1, 1]])


def _ramp(w, h):
    y, x = np.mgrid[0:h, 0:w]
    f = np.stack([x / w * 4.0, y / h * 0.5, np.full_like(x, 0.25, dtype=np.float64)], axis=-1).astype(np.float32)
    f[0, 0] = [100.0, 0.001, 0.0]
    return f


def test_exr_writer(tmp_path):
    import cv2
    f = _ramp(37, 21)     # not a multiple of the 16-line ZIP blocks
    p = str(tmp_path / "out" / "a.exr")
    capi.save_image(p, f)  # also creates the directory (basic.hpp:510-521)
    img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert img is not None and img.shape == (21, 37, 3) and img.dtype == np.float32
    # cv2 returns BGR, top row first; the film's row 0 is the bottom scanline (basic.hpp:583-589)
    assert np.array_equal(img[::-1, :, ::-1], f)
    raw = open(p, "rb").read()
    assert raw[:4] == bytes([0x76, 0x2F, 0x31, 0x01]) and b"channels\x00chlist\x00" in raw
    i = raw.index(b"chlist\x00") + 7 + 4
    assert raw[i:i + 2] == b"B\x00" and struct.unpack("<i", raw[i + 2:i + 6])[0] == 2     # B first, FLOAT


def test_hdr_writer(tmp_path):
    import cv2
    f = _ramp(32, 16)
    p = str(tmp_path / "a.hdr")
    capi.save_image(p, f)
    img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert img is not None and img.shape == (16, 32, 3)
    got = img[::-1, :, ::-1]
    # RGBE keeps 8 bits of mantissa relative to the pixel's largest channel
    mx = np.maximum(f.max(axis=2, keepdims=True), 1e-30)
    assert np.all(np.abs(got - f) <= mx / 128.0 + 1e-6)


def test_png_writer(tmp_path):
    import cv2
    f = _ramp(20, 10)
    p = str(tmp_path / "a.png")
    capi.save_image(p, f)
    img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert img.shape == (10, 20, 3) and img.dtype == np.uint8
    expect = np.clip((np.power(f.astype(np.float64), 1 / 2.2) * 255.0).astype(np.int64), 0, 255).astype(np.uint8)   # basic.hpp:633-646
    assert np.array_equal(img[::-1, :, ::-1], expect)
    with pytest.raises(capi.NgiError):
        capi.save_image(str(tmp_path / "a.bmp"), f)


# ---- TexR textures: image readers + loader ---------------------------------------------------------
def _png_bytes(img8, filters):
    """Minimal PNG encoder applying the given scanline filter type per row (exercises all five unfilter paths)."""
    import struct
    import zlib
    h, w, ch = img8.shape
    raw = bytearray()
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        cur = img8[y].reshape(-1).astype(np.int32)
        ft = filters[y % len(filters)]
        left = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        ul = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - left
        elif ft == 2:
            f = cur - prev
        elif ft == 3:
            f = cur - (left + prev) // 2
        else:
            p = left + prev - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            f = cur - pred
        raw.append(ft)
        raw += bytes((f % 256).astype(np.uint8))
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b"")


@pytest.mark.parametrize("ch", [1, 3, 4])
def test_png_reader_all_filters(tmp_path, ch):
    rng = np.random.default_rng(ch)
    img = rng.integers(0, 256, (13, 17, ch), dtype=np.uint8)
    p = tmp_path / "t.png"
    p.write_bytes(_png_bytes(img, [0, 1, 2, 3, 4]))
    got = capi.load_image(str(p))
    want = (img[..., :3] if ch >= 3 else np.repeat(img[..., :1], 3, axis=2)).astype(np.float32) / 255.0   # rt.hpp:245-250
    assert got.shape == (13, 17, 3) and np.array_equal(got, want)


def test_hdr_and_pfm_readers_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    film = (rng.random((9, 40, 3)) * 4).astype(np.float32)          # row 0 = bottom (film convention)
    capi.save_image(str(tmp_path / "f.hdr"), film)
    got = capi.load_image(str(tmp_path / "f.hdr"))                   # row 0 = top (texture convention)
    assert got.shape == film.shape
    assert np.abs(got - film[::-1]).max() <= 4.0 / 256   # RGBE: 8-bit mantissas under the pixel's largest exponent
    from nanogi_b200 import scenes
    scenes.write_pfm(str(tmp_path / "f.pfm"), film)
    assert np.array_equal(capi.load_image(str(tmp_path / "f.pfm")), film)
    with pytest.raises(capi.NgiError):
        (tmp_path / "x.bin").write_bytes(b"not an image at all")
        capi.load_image(str(tmp_path / "x.bin"))


def test_scene_with_texr_loads_through_the_front_end(tmp_path):
    from nanogi_b200 import scenes
    spec = scenes.cornell_textured()
    path = scenes.write_scene_files(spec, str(tmp_path))
    sd, ref = capi.load_scene_file(path, 1.0), scenes.to_scene_data(spec, 1.0)
    assert len(sd.textures) == 2 and all(np.array_equal(a, b) for a, b in zip(sd.textures, ref.textures))
    assert np.array_equal(sd.texcoords, ref.texcoords) and np.array_equal(sd.positions, ref.positions)
    assert [(p.d_tex, p.g_tex) for p in sd.prims] == [(p.d_tex, p.g_tex) for p in ref.prims]
    assert sum(p.d_tex >= 0 for p in sd.prims) == 1 and sum(p.g_tex >= 0 for p in sd.prims) == 1


def test_cli_b200_additive_options_and_renderer_names():
    o = capi.parse_cli(["nanogi", "ltdirect", "s.yml", "o.pfm", "64", "64", "-t", "2.5", "--progress-image-update-interval", "0.5",
                        "--progress-image-update-format", "p/{{count}}.hdr", "--sample-offset", "4096", "--resume-from", "prev.pfm"])
    assert o.renderer == b"ltdirect" and o.render_time == 2.5 and o.progress_image_update_interval == 0.5
    assert o.progress_image_update_format == b"p/{{count}}.hdr" and o.sample_offset == 4096 and o.resume_from == b"prev.pfm"


def test_pfm_film_round_trip(tmp_path):
    """The additive lossless film format behind --resume-from: written bottom-up like the film, read back top-down."""
    rng = np.random.default_rng(3)
    film = rng.random((5, 7, 3), dtype=np.float32) * 100
    path = str(tmp_path / "f.pfm")
    capi.save_image(path, film)
    img = capi.load_image(path)
    assert np.array_equal(img[::-1], film)


def test_area_sensor_scene_round_trips_through_yaml(tmp_path):
    """E.area sensors: `type: area` + `We`, needs a mesh with uv (rt.hpp:1908-1928)."""
    spec = scenes.cornell_raw_sensor()
    path = scenes.write_scene_files(spec, str(tmp_path))
    sd_file, sd_mem = capi.load_scene_file(path, 1.0), scenes.to_scene_data(spec, 1.0)
    assert sd_file.texcoords is not None and np.allclose(sd_file.texcoords, sd_mem.texcoords)
    a, b = sd_file.prims[sd_file.sensor_prim()], sd_mem.prims[sd_mem.sensor_prim()]
    assert a.e_type == b.e_type == capi.E_AREA and list(a.e_we) == [1.0, 1.0, 1.0] and a.num_tris == 2
    from oracle import pyoracle
    fa, _ = pyoracle.OracleScene(sd_file).render("lt", 2000, 8, 8, max_num_vertices=4, seed=1, rng_mode=1)
    fb, _ = pyoracle.OracleScene(sd_mem).render("lt", 2000, 8, 8, max_num_vertices=4, seed=1, rng_mode=1)
    assert np.array_equal(fa, fb) and fa.max() > 0
