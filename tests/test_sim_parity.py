"""Device algorithms vs the oracle, stepped on the CPU by the test-only simulator (tests/hostsim). These are
the same checks tests/test_gpu_parity.py runs on the B200 through the C ABI."""
import numpy as np
import pytest

from nanogi_b200 import capi, scenes
from oracle import pyoracle
from tests import parity_common as pc
from tests.conftest import scaled_spec
from tests.hostsim import pysim


def test_device_philox_matches_known_answers():
    assert [hex(x) for x in pysim.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    rng = np.random.default_rng(0)
    for _ in range(50):
        c, k = rng.integers(0, 2 ** 32, 4, dtype=np.uint64), rng.integers(0, 2 ** 32, 2, dtype=np.uint64)
        assert np.array_equal(pysim.philox(c, k), pyoracle.philox(c, k))


@pytest.mark.parametrize("name", ["cornell", "cornell_spheres", "furnace"])
def test_trace_bit_exact(name, request):
    sd = request.getfixturevalue(name)
    pc.check_trace_bit_exact(pysim.SimScene(sd), sd, n_random=8000, cam=48)


def test_trace_edge_cases(cornell, cornell_spheres):
    pc.check_trace_edge_cases(pysim.SimScene(cornell), cornell)
    pc.check_trace_edge_cases(pysim.SimScene(cornell_spheres), cornell_spheres)


def test_trace_bit_exact_medium_mesh():
    """~20k triangles: 1 subdiv-4 icosphere per material + room; oracle BVH cross-checked by brute force on a subset."""
    spec = scenes.instanced_spheres(seed=1, subdiv=3, grid=(2, 2, 2))
    sd = scenes.to_scene_data(spec, 16 / 9)
    sim = pysim.SimScene(sd)
    info = sim.info()
    assert info["n"] == sd.num_tris and info["nodes8"] < info["n"] and info["depth8"] <= 40
    pc.check_trace_bit_exact(sim, sd, n_random=6000, cam=32, accels=(0, 1))
    orc = pyoracle.OracleScene(sd)
    rays = scenes.random_rays(sd, 300, 9)
    assert np.array_equal(orc.trace(rays, 0), orc.trace(rays, 2))


def test_degenerate_scenes():
    """empty geometry, a single triangle, coincident duplicates (Morton ties)."""
    cam = scenes.pinhole(eye=[0, 0, 5], center=[0, 0, 0], up=[0, 1, 0], fov_deg=40)
    tri = np.array([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]], dtype=np.float64)
    light = scenes.mesh_prim(["L", "D"], tri, name="l", L={"type": "area", "Le": [1, 1, 1]}, D={"R": [0, 0, 0]})
    for spec in ([light, cam], [scenes.mesh_prim(["D"], np.repeat(tri, 9, axis=0), name="dup", D={"R": [1, 1, 1]}), light, cam]):
        sd = scenes.to_scene_data(spec, 1.0)
        sim = pysim.SimScene(sd)
        pc.check_trace_bit_exact(sim, sd, n_random=2000, cam=16)
    # no mesh at all: every ray misses
    pt_light = {"type": ["L"], "mesh": None, "params": {"L": {"type": "point", "Le": [1, 1, 1], "position": [0, 1, 0]}}}
    sd = scenes.to_scene_data([pt_light, cam], 1.0)
    sim = pysim.SimScene(sd)
    rays = scenes.camera_rays(sd, 8, 8)
    assert (sim.trace(rays, False, 0)["tri"] == capi.NO_HIT).all()
    film, st = sim.render("ptdirect", 1000, 8, 8, seed=1)
    assert st["extend_rays"] == 1000 and np.isfinite(film).all()


@pytest.mark.parametrize("prim,type_bit", [(4, capi.TYPE_D), (8, capi.TYPE_G), (9, capi.TYPE_S)])
def test_bsdf_parity(cornell_spheres, prim, type_bit):
    pc.check_bsdf_parity(pysim.SimScene(cornell_spheres), cornell_spheres, prim, type_bit, n=1500)


def test_bsdf_parity_reflection_refraction():
    cam = scenes.pinhole(eye=[0, 0, 5], center=[0, 0, 0], up=[0, 1, 0], fov_deg=40)
    tri = np.array([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]], dtype=np.float64)
    spec = [scenes.mesh_prim(["S"], tri, name="mirror", S={"type": "reflection", "R": [0.9, 0.8, 0.7]}),
            scenes.mesh_prim(["S"], tri + 2, name="glass", S={"type": "refraction", "R": [1, 1, 1], "eta1": 1.0, "eta2": 1.5}), cam]
    sd = scenes.to_scene_data(spec, 1.0)
    sim = pysim.SimScene(sd)
    pc.check_bsdf_parity(sim, sd, 0, capi.TYPE_S, n=800)
    pc.check_bsdf_parity(sim, sd, 1, capi.TYPE_S, n=800, seed=1)


@pytest.fixture(scope="module")
def branches_sim():
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_branches(), 0.01), 1.0)
    return pysim.SimScene(sd), sd


@pytest.mark.parametrize("name,bits", [("dg", capi.TYPE_D | capi.TYPE_G), ("gs", capi.TYPE_G | capi.TYPE_S), ("refr", capi.TYPE_S), ("mirror", capi.TYPE_S)])
def test_bsdf_parity_by_precedence(branches_sim, name, bits):
    """CPU twin of tests/test_gpu_branches.py: [D, G] evaluates as D, [G, S] as G (rt.hpp:808-859)"""
    sim, sd = branches_sim
    prim = pc.prim_index(scenes.cornell_branches(), name)
    assert sd.prims[prim].type & capi.TYPE_BSDF_MASK == bits
    pc.check_bsdf_parity(sim, sd, prim, bits, n=600, seed=3)


@pytest.mark.parametrize("renderer", ["pt", "ptdirect", "ltdirect"])
def test_replay_branches(branches_sim, renderer):
    """S.reflection / S.refraction, two-lobe primitives, a pure [L] mesh: sample-exact against the oracle"""
    sim, sd = branches_sim
    pc.check_replay(sim, sd, renderer, n=20000, m=-1, wave_capacity=2048, max_bad_pixels=0.006)


def test_replay_large_light_cdf():
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_branches(light_res=64), 0.01), 1.0)      # 8 192-triangle light
    pc.check_replay(pysim.SimScene(sd), sd, "ptdirect", n=20000, m=4, wave_capacity=2048, max_bad_pixels=0.006)


@pytest.mark.parametrize("renderer", ["pt", "ptdirect"])
@pytest.mark.parametrize("m", [-1, 3])
def test_replay_small_scale(renderer, m):
    """C2's scene scaled to unit size (outside the fp32 self-intersection regime): sample-exact agreement."""
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_spheres(), 0.01), 1.0)
    pc.check_replay(pysim.SimScene(sd), sd, renderer, n=20000, m=m, wave_capacity=2048)


@pytest.mark.parametrize("renderer", ["pt", "ptdirect"])
def test_replay_textured(renderer):
    """D.TexR / G.TexR (nearest texel, fract wrap, uv interpolated at the hit) against the oracle's Texture."""
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_textured(), 0.01), 1.0)
    pc.check_replay(pysim.SimScene(sd), sd, renderer, n=20000, m=-1, wave_capacity=2048)


# ---- lt / ltdirect, E.area sensors, point + directional lights -------------------------
@pytest.mark.parametrize("renderer", ["lt", "ltdirect"])
@pytest.mark.parametrize("m", [-1, 3])
def test_replay_light_tracing(renderer, m):
    """Light paths (TransportDirection::LE: shading-normal correction ratio, no (eta_i/eta_t)^2) through D, G and S-fresnel
    lobes; ltdirect connects every vertex to the pinhole. `lt` alone can only hit a sensor that has a mesh: with the
    pinhole its film is exactly zero on both sides while the rays are still traced."""
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_spheres(), 0.01), 1.0)
    pc.check_replay(pysim.SimScene(sd), sd, renderer, n=20000, m=m, wave_capacity=2048)


@pytest.mark.parametrize("renderer", ["pt", "ptdirect", "lt", "ltdirect"])
def test_replay_area_sensor(renderer):
    """E.area ("raw") sensor: position sampled on the sensor mesh, raster position = uv, We iff cos > 0; `lt` splats
    on hits of the sensor mesh; ltdirect keeps the reference's `L->EvaluatePositionPDF(geomE)` quirk."""
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_raw_sensor(spheres=True), 0.01), 1.0)
    pc.check_replay(pysim.SimScene(sd), sd, renderer, n=20000, w=16, h=16, m=6, wave_capacity=2048)


@pytest.mark.parametrize("renderer", ["ptdirect", "lt", "ltdirect"])
def test_replay_point_and_directional_lights(renderer):
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_mixed_lights(), 0.01), 1.0)
    pc.check_replay(pysim.SimScene(sd), sd, renderer, n=20000, m=5, wave_capacity=2048)


def test_light_tracing_statistics_cornell_scale(cornell):
    """ltdirect at Cornell scale (fp32 self-intersection regime): statistical agreement with the oracle."""
    pc.check_image_statistics(pysim.SimScene(cornell), cornell, "ltdirect", w=16, h=16, spp=256, seeds=6, m=6, block=4, wave_capacity=4096)


@pytest.mark.parametrize("scene,m", [("cornell_spheres", -1), ("cornell_mixed_lights", 5), ("cornell_raw_sensor", 6), ("cornell_textured", 4)])
def test_replay_bdpt(scene, m):
    """bdpt (one sample per thread, ngi_bdpt.h) against the oracle's restatement of bdpt.hpp — itself bit-exact against the
    reference's own code (tests/test_reference_pin.py): same Philox uniforms, sample-exact films and ray counts at unit scale."""
    spec = scenes.cornell_raw_sensor(spheres=True) if scene == "cornell_raw_sensor" else getattr(scenes, scene)()
    sd = scenes.to_scene_data(scaled_spec(spec, 0.01), 1.0)
    pc.check_replay(pysim.SimScene(sd), sd, "bdpt", n=15000, w=24, h=24, m=m)


@pytest.mark.parametrize("scene,m,batch", [("cornell_spheres", -1, 1000), ("cornell_raw_sensor", 6, 4096), ("cornell_mixed_lights", 3, 333), ("cornell_textured", 4, 2048)])
def test_bdpt_wavefront_equals_per_thread(scene, m, batch):
    """The wavefront stages (ngi_bdpt_wave.h: batches of samples through start / extend / step / count / expand / shadow /
    contrib) and the one-sample-per-thread form (ngi_bdpt.h) run the same functions on the same Philox counters: identical
    ray counts, and films equal up to the order of the float additions — for any batch size (ragged last batch included)."""
    spec = scenes.cornell_raw_sensor(spheres=True) if scene == "cornell_raw_sensor" else getattr(scenes, scene)()
    sd = scenes.to_scene_data(scaled_spec(spec, 0.01), 1.0)
    sim = pysim.SimScene(sd)
    fa, sa = sim.render("bdpt", 10000, 24, 24, max_num_vertices=m, seed=5, sample_offset=77, wave_capacity=batch)
    fb, sb = sim.render("bdpt", 10000, 24, 24, max_num_vertices=m, seed=5, sample_offset=77, flags=capi.RENDER_BDPT_PER_THREAD)
    assert sa["extend_rays"] == sb["extend_rays"] and sa["shadow_rays"] == sb["shadow_rays"]
    assert fa.sum() > 0
    np.testing.assert_allclose(fa, fb, rtol=2e-5, atol=1e-6 * float(fb.max()))


@pytest.mark.parametrize("n", [1, 7, 1000])
@pytest.mark.parametrize("m", [2, 3, -1])
def test_bdpt_wavefront_edge_sizes(n, m):
    """Edge sizes of the batch machinery: fewer samples than a batch, batches of 3 samples (ragged last batch), the shortest paths."""
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_spheres(), 0.01), 1.0)
    sim = pysim.SimScene(sd)
    fb, sb = sim.render("bdpt", n, 16, 16, max_num_vertices=m, seed=11, sample_offset=5, film_norm_samples=1000, flags=capi.RENDER_BDPT_PER_THREAD)
    for batch in (3, 4096):
        fa, sa = sim.render("bdpt", n, 16, 16, max_num_vertices=m, seed=11, sample_offset=5, film_norm_samples=1000, wave_capacity=batch)
        assert sa["extend_rays"] == sb["extend_rays"] and sa["shadow_rays"] == sb["shadow_rays"]
        np.testing.assert_allclose(fa, fb, rtol=2e-5, atol=1e-6 * max(float(fb.max()), 1e-30))


def test_bdpt_wavefront_without_lights():
    """No light primitive: the light subpaths are empty, every sample ends after the eye subpath was traced (src/nanogi.cpp:1137-1146), the
    film stays black — in both forms."""
    spec = [p for p in scaled_spec(scenes.cornell_box(), 0.01) if "L" not in p["type"]]
    sd = scenes.to_scene_data(spec, 1.0)
    sim = pysim.SimScene(sd)
    fa, sa = sim.render("bdpt", 2000, 16, 16, max_num_vertices=5, seed=3, wave_capacity=512)
    fb, sb = sim.render("bdpt", 2000, 16, 16, max_num_vertices=5, seed=3, flags=capi.RENDER_BDPT_PER_THREAD)
    assert not fa.any() and not fb.any()
    assert sa["shadow_rays"] == sb["shadow_rays"] == 0 and sa["extend_rays"] == sb["extend_rays"] > 0


def test_bdpt_statistics_cornell_scale(cornell):
    pc.check_image_statistics(pysim.SimScene(cornell), cornell, "bdpt", w=16, h=16, spp=128, seeds=6, m=6, block=4)


def test_replay_furnace_exact(furnace):
    sim = pysim.SimScene(furnace)
    orc = pyoracle.OracleScene(furnace)
    fo, so = orc.render("pt", 30000, 16, 16, seed=9, rng_mode=1)
    fs, ss = sim.render("pt", 30000, 16, 16, seed=9, wave_capacity=1024)
    assert ss["extend_rays"] == so["extend_rays"]
    assert np.allclose(fs, fo, rtol=1e-5)


def test_wave_capacity_does_not_change_the_sample_set(cornell):
    sim = pysim.SimScene(cornell)
    a, sa = sim.render("ptdirect", 5000, 16, 16, seed=4, wave_capacity=64)
    b, sb = sim.render("ptdirect", 5000, 16, 16, seed=4, wave_capacity=5000)
    assert sa["extend_rays"] == sb["extend_rays"] and sa["shadow_rays"] == sb["shadow_rays"]
    assert np.allclose(a, b, rtol=1e-4, atol=1e-6)


def test_sharding(cornell):
    pc.check_sharding(pysim.SimScene(cornell), wave_capacity=1024)


def test_self_intersection_rate_matches_reference_arithmetic():
    """The reference rebuilds the hit point along the fp64 direction whose fp32 rounding was traced
    (rt.hpp:2169-2171 vs :2197); with the absolute 1e-4 epsilon that decides how often a bounce ray re-hits its
    own wall at Cornell scale. pt with -m 3 in a furnace of side 500 at offset 500 measures the rate a:
    E = 1 + rho (1 - a). The device path must reproduce the oracle's a (it would be ~6 % lower without the
    direction dither of ngi_wave.h)."""
    sd = scenes.to_scene_data(scaled_spec(scenes.furnace(0.5, 1.0), 250.0, 500.0), 1.0)
    orc, sim = pyoracle.OracleScene(sd), pysim.SimScene(sd)
    n = 1 << 21
    fo, _ = orc.render("pt", n, 8, 8, max_num_vertices=3, seed=1)
    fs, _ = sim.render("pt", n, 8, 8, max_num_vertices=3, seed=2, wave_capacity=1 << 14)
    a_o, a_s = (1.5 - fo.mean()) * 2, (1.5 - fs.mean()) * 2
    assert 0.005 < a_o < 0.05
    assert abs(a_s - a_o) < 0.0015, (a_s, a_o)


@pytest.mark.parametrize("renderer", ["pt", "ptdirect"])
def test_image_statistics_cornell(cornell, renderer):
    pc.check_image_statistics(pysim.SimScene(cornell), cornell, renderer, w=16, h=16, spp=256 if renderer == "ptdirect" else 1024,
                              seeds=6, m=8, block=4, wave_capacity=4096)


@pytest.mark.parametrize("renderer,m,expect", [("pt", -1, 2.0), ("pt", 2, 1.0), ("pt", 4, 1.75), ("ptdirect", 3, 1.5)])
def test_furnace(renderer, m, expect):
    pc.check_furnace(lambda sd: pysim.SimScene(sd), renderer, m, expect, n=1 << 17, tol=0.02 if renderer == "pt" else 0.05)
