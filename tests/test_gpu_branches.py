"""GPU parity for the branches no BASELINE scene takes (VERDICT r1 "untested branches"): S.reflection / S.refraction
(rt.hpp:808-859, :1066-1098), primitives with two BSDF types resolved by precedence (rt.hpp:338-351, schema.yml:60-62), pure [L]
meshes (rt.hpp:909, :1146, :1334), a light with a large area CDF (basic.hpp:440-497, rt.hpp:1747-1765) and the direction dither at
C3 / C4 coordinate scales. The oracle these compare with is pinned bit for bit against the reference's own code on the same scene
(tests/test_reference_pin.py: "cornell_branches"). CPU-simulator twins: tests/test_sim_parity.py."""
import numpy as np
import pytest

from nanogi_b200 import capi, scenes
from oracle import pyoracle
from tests import parity_common as pc
from tests.conftest import scaled_spec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def branches():
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_branches(), 0.01), 1.0)
    g = capi.GpuScene(sd, 0)
    yield g, sd
    g.close()


def test_bsdf_parity_reflection_refraction():
    cam = scenes.pinhole(eye=[0, 0, 5], center=[0, 0, 0], up=[0, 1, 0], fov_deg=40)
    tri = np.array([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]], dtype=np.float64)
    spec = [scenes.mesh_prim(["S"], tri, name="mirror", S={"type": "reflection", "R": [0.9, 0.8, 0.7]}),
            scenes.mesh_prim(["S"], tri + 2, name="glass", S={"type": "refraction", "R": [1, 1, 1], "eta1": 1.0, "eta2": 1.5}), cam]
    sd = scenes.to_scene_data(spec, 1.0)
    g = capi.GpuScene(sd, 0)
    pc.check_bsdf_parity(g, sd, 0, capi.TYPE_S, n=3000)
    pc.check_bsdf_parity(g, sd, 1, capi.TYPE_S, n=3000, seed=1)
    g.close()


@pytest.mark.parametrize("name,bits", [("dg", capi.TYPE_D | capi.TYPE_G), ("gs", capi.TYPE_G | capi.TYPE_S), ("refr", capi.TYPE_S), ("mirror", capi.TYPE_S)])
def test_bsdf_parity_by_precedence(branches, name, bits):
    """[D, G] evaluates as D, [G, S] as G: queried with the primitive's full type mask, like the path loop does (src/nanogi.cpp:601)"""
    g, sd = branches
    prim = pc.prim_index(scenes.cornell_branches(), name)
    assert sd.prims[prim].type & capi.TYPE_BSDF_MASK == bits
    pc.check_bsdf_parity(g, sd, prim, bits, n=2500, seed=3)


@pytest.mark.parametrize("renderer", ["pt", "ptdirect", "ltdirect"])
@pytest.mark.parametrize("m", [-1, 4])
def test_replay_branches(branches, renderer, m):
    g, sd = branches
    pc.check_replay(g, sd, renderer, n=200000, w=96, h=96, m=m, max_bad_pixels=0.012)


def test_bare_light_is_seen_and_ends_paths(branches):
    """the [L]-only quad: `pt` adds its emission at the hit (src/nanogi.cpp:566-577) and a path that hits it has no lobe to continue
    with. Same Philox uniforms on both sides (the oracle's counter-based mode), so film sum and rays per path agree closely; the
    pixels that see the quad directly are brighter than their neighbourhood on both sides."""
    g, sd = branches
    orc = pyoracle.OracleScene(sd)
    n = 1 << 20
    fg, sg = g.render("pt", n, 64, 64, seed=4)
    fo, so = orc.render("pt", n, 64, 64, seed=4, rng_mode=1)
    assert abs(float(fg.sum(dtype=np.float64)) - fo.sum()) < 0.01 * fo.sum()
    assert abs(sg.extend_rays - so["extend_rays"]) <= 2e-4 * so["extend_rays"]
    # only the bare quad as light: every contribution of `pt` is a hit of it, and no path continues from it
    spec = [p for p in scaled_spec(scenes.cornell_branches(), 0.01) if not (p.get("mesh") and p["mesh"]["name"] == "light")]
    sd1 = scenes.to_scene_data(spec, 1.0)
    g1, o1 = capi.GpuScene(sd1, 0), pyoracle.OracleScene(sd1)
    a, sa = g1.render("pt", n, 64, 64, seed=9)
    b, sb = o1.render("pt", n, 64, 64, seed=9, rng_mode=1)
    assert b.sum() > 0 and abs(float(a.sum(dtype=np.float64)) - b.sum()) < 0.01 * b.sum()
    assert abs(sa.extend_rays - sb["extend_rays"]) <= 2e-4 * sb["extend_rays"]
    g1.close()


def test_replay_large_light_cdf():
    """131 072-triangle ceiling light: the device's fp32 CDF picks the same triangle as the reference's fp64 one for all but the
    samples that land within rounding of a CDF step"""
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_branches(light_res=256), 0.01), 1.0)
    assert sd.num_tris > 131072
    g = capi.GpuScene(sd, 0)
    pc.check_replay(g, sd, "ptdirect", n=200000, w=96, h=96, m=4, max_bad_pixels=0.02)
    # statistics at full scale too (Cornell coordinates): means agree
    sd2 = scenes.to_scene_data(scenes.cornell_branches(light_res=256), 1.0)
    g2 = capi.GpuScene(sd2, 0)
    fg, _ = g2.render("ptdirect", 1 << 21, 32, 32, seed=8)
    fo, _ = pyoracle.OracleScene(sd2).render("ptdirect", 1 << 18, 32, 32, seed=9)
    assert abs(fg.mean() - fo.mean()) < 0.02 * fo.mean()
    g.close(); g2.close()


@pytest.mark.parametrize("scale,offset", [(10.0, 0.0), (30.0, 0.0), (1.3, 4.0)])
def test_self_intersection_rate_at_c3_c4_coordinate_scales(scale, offset):
    """the hashed +-ulp/2 direction dither (ngi_dither_direction) at the coordinate ranges of C3 (room +-10, spheres of radius ~1 around
    y = 4) and C4 (+-30): the rate of bounce rays that re-hit their own surface equals the oracle's (reference arithmetic:
    fp64 direction traced in fp32, rt.hpp:2169-2171, :2197)"""
    sd = scenes.to_scene_data(scaled_spec(scenes.furnace(0.5, 1.0), scale, offset), 1.0)
    orc, g = pyoracle.OracleScene(sd), capi.GpuScene(sd, 0)
    fo, _ = orc.render("pt", 1 << 22, 8, 8, max_num_vertices=3, seed=1)
    fg, _ = g.render("pt", 1 << 24, 8, 8, max_num_vertices=3, seed=2)
    a_o, a_g = (1.5 - fo.mean()) * 2, (1.5 - float(fg.mean())) * 2
    assert abs(a_g - a_o) < 0.001, (a_g, a_o)
    g.close()
