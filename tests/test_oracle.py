"""The oracle against known answers we derived analytically — the second, independent anchor next to the pin against the
reference's own code (tests/test_reference_pin.py, tests/golden/reference_films.npz). The reference ships no
tests or golden vectors of its own."""
import math

import numpy as np
import pytest

from nanogi_b200 import capi, scenes
from oracle import pyoracle


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in pyoracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in pyoracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in pyoracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_orthonormal_basis():
    rng = np.random.default_rng(0)
    for a in list(rng.normal(size=(20, 3))) + [np.array([0, 0, 1.0]), np.array([0, 1.0, 0]), np.array([1.0, 0, 0]), np.array([0, 0, -1.0])]:
        a = a / np.linalg.norm(a)
        b, c = pyoracle.orthonormal_basis(a)
        assert abs(np.dot(a, b)) < 1e-12 and abs(np.dot(a, c)) < 1e-12 and abs(np.dot(b, c)) < 1e-12
        assert abs(np.linalg.norm(b) - 1) < 1e-12 and abs(np.linalg.norm(c) - 1) < 1e-12
        # ToWorld = [b c a] is right handed: b x c = a
        assert np.allclose(np.cross(b, c), a, atol=1e-12)


def test_fresnel_normal_incidence(cornell_spheres):
    o = pyoracle.OracleScene(cornell_spheres)
    # ((eta1 - eta2)/(eta1 + eta2))^2 : 1 -> 2 gives 1/9
    assert abs(o.fresnel(9, 1.0, 1.0, 2.0) - 1.0 / 9.0) < 1e-15
    assert abs(o.fresnel(9, 1.0, 1.0, 1.5) - 0.04) < 1e-15
    # total internal reflection from the dense side beyond the critical angle
    assert o.fresnel(9, 0.3, 2.0, 1.0) == 1.0
    # grazing incidence -> 1
    assert abs(o.fresnel(9, 1e-9, 1.0, 1.5) - 1.0) < 1e-6


def test_geometry_term():
    g = pyoracle.geometry_term([0, 0, 0], [0, 1, 0], 0, [0, 2, 0], [0, -1, 0], 0)
    assert abs(g - 0.25) < 1e-15
    # degenerate endpoints drop their cosine (rt.hpp:2371-2372)
    g = pyoracle.geometry_term([0, 0, 0], [0, 1, 0], 1, [1, 1, 0], [0, -1, 0], 0)
    assert abs(g - (math.sqrt(0.5) / 2.0)) < 1e-15


def test_light_cdf_and_inv_area(cornell):
    o = pyoracle.OracleScene(cornell)
    cdf, inv_area = o.light_cdf(0)
    assert len(cdf) == 3 and cdf[0] == 0 and abs(cdf[-1] - 1) < 1e-15 and abs(cdf[1] - 0.5) < 1e-12
    assert abs(1 / inv_area - 130 * 105) < 1e-6
    info = o.info()
    assert info["tris"] == 38 and info["lights"] == 1 and info["sensor"] == 8


def test_sample_position_on_light(cornell):
    o = pyoracle.OracleScene(cornell)
    rng = np.random.default_rng(1)
    for u in rng.random((200, 2)):
        s = o.sample_position(0, u[0], u[1])
        assert abs(s["p"][1] - 548.75) < 1e-3 and 213 - 1e-3 <= s["p"][0] <= 343 + 1e-3 and 227 - 1e-3 <= s["p"][2] <= 332 + 1e-3
        assert np.allclose(s["gn"], [0, -1, 0], atol=1e-9) and np.allclose(s["sn"], s["gn"])   # face normal, rt.hpp:519-523
        assert abs(s["pdf"] - 1 / (130 * 105)) < 1e-12


def test_pinhole_roundtrip(cornell):
    """RasterPosition(SampleDirection(u)) == u (SURVEY §8c known-answer 4) and We/pdf == 1."""
    o = pyoracle.OracleScene(cornell)
    E = cornell.sensor_prim()
    rng = np.random.default_rng(2)
    for u in rng.random((100, 2)):
        wo, wrote = o.sample_direction(E, capi.TYPE_E, [0, 0, 1], [0, 0, 1], [0, 0, 0], u[0], u[1], 0.5)
        assert wrote and abs(np.linalg.norm(wo) - 1) < 1e-12
        ok, rx, ry, pix = o.raster_position(E, wo, 256, 256)
        assert ok and abs(rx - u[0]) < 1e-12 and abs(ry - u[1]) < 1e-12
        assert pix == min(int(u[1] * 256), 255) * 256 + min(int(u[0] * 256), 255)
        fs, pdf = o.evaluate_direction(E, capi.TYPE_E, [0, 0, 1], [0, 0, 1], [0, 0, 0], wo)
        assert pdf > 0 and abs(fs[0] / pdf - 1) < 1e-12
    ok, *_ = o.raster_position(E, [0, 0, -1.0], 8, 8)   # behind the camera (camera looks along +z here)
    assert not ok


def test_diffuse_bsdf_values(cornell):
    o = pyoracle.OracleScene(cornell)
    sn = gn = [0, 1, 0]
    wi = np.array([0.3, 0.8, 0.1]); wi /= np.linalg.norm(wi)
    wo = np.array([-0.2, 0.5, 0.4]); wo /= np.linalg.norm(wo)
    fs, pdf = o.evaluate_direction(4, capi.TYPE_D, sn, gn, wi, wo)   # red wall: R = (1, 0, 0)
    assert np.allclose(fs, [1 / math.pi, 0, 0]) and abs(pdf - 1 / math.pi) < 1e-15
    fs, pdf = o.evaluate_direction(4, capi.TYPE_D, sn, gn, wi, -wo)  # below the surface
    assert np.all(fs == 0) and pdf == 0
    # shading-normal correction: geometric side disagrees with the shading side -> 0
    fs, _ = o.evaluate_direction(4, capi.TYPE_D, sn, [0, -1, 0], wi, wo)
    assert np.all(fs == 0)
    # cosine sampling never leaves the hemisphere; "wo not written" when wi is below
    wo2, wrote = o.sample_direction(4, capi.TYPE_D, sn, gn, -wi, 0.3, 0.6, 0.5)
    assert not wrote and np.all(wo2 == 0)


def test_energy_of_diffuse_lobe(cornell):
    """int fs cos = rho <= 1 for D when integrated with its own sampler (fs/pdf in projected solid angle)."""
    o = pyoracle.OracleScene(cornell)
    rng = np.random.default_rng(3)
    acc = 0.0
    wi = np.array([0.0, 1.0, 0.0])
    n = 2000
    for u in rng.random((n, 2)):
        wo, wrote = o.sample_direction(1, capi.TYPE_D, [0, 1, 0], [0, 1, 0], wi, u[0], u[1], 0.5)
        fs, pdf = o.evaluate_direction(1, capi.TYPE_D, [0, 1, 0], [0, 1, 0], wi, wo)
        acc += fs[0] / pdf
    assert abs(acc / n - 1.0) < 1e-9


def test_glossy_matches_reference_formula_not_physics(cornell_spheres):
    """The replicated shadow-masking typo: G uses |wo.H| in both terms (rt.hpp:1428-1429)."""
    o = pyoracle.OracleScene(cornell_spheres)
    sn = gn = np.array([0, 0, 1.0])
    wi = np.array([0.6, 0.0, 0.8])
    wo = np.array([-0.5, 0.1, 0.86]); wo /= np.linalg.norm(wo)
    fs, pdf = o.evaluate_direction(8, capi.TYPE_G, sn, gn, wi, wo)
    b, c = pyoracle.orthonormal_basis(sn)
    lwi = np.array([b @ wi, c @ wi, sn @ wi]); lwo = np.array([b @ wo, c @ wo, sn @ wo])
    H = lwi + lwo; H /= np.linalg.norm(H)
    a = 0.1
    tan2 = (1 - H[2] ** 2) / H[2] ** 2
    D = math.exp(-tan2 / a ** 2) / (math.pi * a ** 2 * H[2] ** 4)
    G = min(1.0, 2 * H[2] * lwo[2] / abs(lwo @ H), 2 * H[2] * lwi[2] / abs(lwo @ H))
    eta = np.array(scenes.COPPER["Eta"]); k = np.array(scenes.COPPER["K"]); R = np.array(scenes.COPPER["R"])
    ci = lwi @ H
    tmp = (eta ** 2 + k ** 2) * ci ** 2
    rpar = (tmp - 2 * eta * ci + 1) / (tmp + 2 * eta * ci + 1)
    tf = eta ** 2 + k ** 2
    rper = (tf - 2 * eta * ci + ci ** 2) / (tf + 2 * eta * ci + ci ** 2)
    F = (rpar + rper) / 2
    expect = R * D * G * F / (4 * lwi[2]) / lwo[2]
    assert np.allclose(fs, expect, rtol=1e-12)
    assert abs(pdf - D * H[2] / (4 * (lwo @ H)) / lwo[2]) < 1e-12 * abs(pdf)


def test_fresnel_sampling_branches(cornell_spheres):
    o = pyoracle.OracleScene(cornell_spheres)
    sn = gn = np.array([0, 0, 1.0])
    wi = np.array([0.6, 0.0, 0.8])
    Fr = o.fresnel(9, 0.8, 1.0, 2.0)
    wo_r, _ = o.sample_direction(9, capi.TYPE_S, sn, gn, wi, 0.1, 0.2, Fr * 0.999)
    b, c = pyoracle.orthonormal_basis(sn)
    lwi = np.array([b @ wi, c @ wi, sn @ wi])
    lwo = np.array([b @ wo_r, c @ wo_r, sn @ wo_r])
    assert np.allclose(lwo, [-lwi[0], -lwi[1], lwi[2]])
    fs, pdf = o.evaluate_direction(9, capi.TYPE_S, sn, gn, wi, wo_r)
    assert abs(pdf - Fr) < 1e-15 and np.allclose(fs / pdf, [0.60784313725, 0.80392156862, 1])
    wo_t, _ = o.sample_direction(9, capi.TYPE_S, sn, gn, wi, 0.1, 0.2, min(1.0, Fr * 1.001 + 1e-9))
    lwt = np.array([b @ wo_t, c @ wo_t, sn @ wo_t])
    assert lwt[2] < 0 and abs(np.linalg.norm(lwt) - 1) < 1e-12
    # Snell: sin_t = sin_i / 2
    assert abs(math.hypot(lwt[0], lwt[1]) - 0.6 / 2.0) < 1e-12
    fs, pdf = o.evaluate_direction(9, capi.TYPE_S, sn, gn, wi, wo_t)
    assert abs(pdf - (1 - Fr)) < 1e-15
    assert np.allclose(fs / pdf, np.array([0.60784313725, 0.80392156862, 1]) * 0.25)   # (eta_i/eta_t)^2, rt.hpp:1130-1132
    # not force-degenerated (NEE): specular evaluates to zero, rt.hpp:1057-1060
    fs, pdf = o.evaluate_direction(9, capi.TYPE_S, sn, gn, wi, wo_r, force_degenerated=False)
    assert np.all(fs == 0) and pdf == 0


@pytest.mark.parametrize("name", ["cornell", "cornell_spheres", "furnace"])
def test_oracle_bvh_is_a_pure_filter(name, request):
    """BVH traversal == O(N) brute force, bit for bit."""
    sd = request.getfixturevalue(name)
    o = pyoracle.OracleScene(sd)
    rays = np.concatenate([scenes.camera_rays(sd, 48, 48), scenes.random_rays(sd, 6000, 11)])
    a, b = o.trace(rays, 0), o.trace(rays, 2)
    assert np.array_equal(a, b)
    occ = scenes.random_rays(sd, 6000, 12, occlusion=True)
    any_hits = o.trace(occ, 1)
    closest = o.trace(occ, 2)
    assert np.array_equal(any_hits["tri"] == 0, closest["tri"] != capi.NO_HIT)


def test_intersect_reconstruction(cornell):
    o = pyoracle.OracleScene(cornell)
    # straight at the back wall through the box centre
    h = o.intersect([100, 450, -800], [0, 0, 1])
    assert h is not None and h["tri"] in (2, 3)
    assert abs(h["p"][2] - 559.2) < 1e-3 and np.allclose(h["gn"], [0, 0, -1], atol=1e-6) and np.allclose(h["sn"], [0, 0, -1], atol=1e-6)
    assert o.intersect([100, 450, -800], [0, 0, -1]) is None
    # Visible: light centre <-> floor centre is unobstructed, floor <-> point above the ceiling is blocked
    assert o.visible([278, 0, 279.5], [278, 548.75, 279.5])
    assert not o.visible([278, 0.0, 279.5], [278, 700.0, 279.5])


@pytest.mark.parametrize("renderer", ["pt", "ptdirect"])
@pytest.mark.parametrize("m,expect", [(-1, 2.0), (2, 1.0), (3, 1.5), (4, 1.75)])
def test_furnace(furnace, renderer, m, expect):
    """Closed box, all walls [L, D] Le = 1, rho = 0.5: sum_{k < M-1} rho^k (SURVEY §8c known-answer 1)."""
    o = pyoracle.OracleScene(furnace)
    n = 1 << 18 if renderer == "pt" else 1 << 20
    film, st = o.render(renderer, n, 8, 8, max_num_vertices=m, seed=123 + m)
    tol = 0.01 if renderer == "pt" else 0.03   # ptdirect has heavy tails in the box corners
    assert abs(film.mean() - expect) < tol * expect
    if m == 2 and renderer == "pt":
        assert np.allclose(film.mean(), 1.0, atol=1e-9)   # every path contributes exactly Le


@pytest.mark.parametrize("renderer", ["pt", "ptdirect"])
def test_light_over_plane_form_factor(renderer):
    """Single area light over a diffuse plane with -m 3: radiance = rho * Le * F (known-answer 2)."""
    le, height, half, rho = 5.0, 1.0, 0.5, 0.8
    sd = scenes.to_scene_data(scenes.light_over_plane(le, height, half, rho), 1.0)
    o = pyoracle.OracleScene(sd)
    X = half / math.sqrt(half * half + height * height)
    expect = rho * le * 4 / math.pi * X * math.atan(X)
    n = 1 << 21 if renderer == "pt" else 1 << 18
    film, _ = o.render(renderer, n, 2, 2, max_num_vertices=3, seed=9)
    assert abs(film.mean() - expect) < 0.015 * expect


def test_pt_equals_ptdirect_in_expectation(cornell):
    """Estimator agreement (known-answer 3): block means of pt and ptdirect agree within Monte-Carlo error."""
    o = pyoracle.OracleScene(cornell)
    w = h = 16
    fa = np.mean([o.render("pt", w * h * 8192, w, h, seed=s)[0] for s in (1, 2, 3, 4)], axis=0)
    fb = np.mean([o.render("ptdirect", w * h * 2048, w, h, seed=s)[0] for s in (5, 6, 7, 8)], axis=0)
    assert abs(fa.mean() - fb.mean()) < 0.04 * fb.mean()   # both estimators are noisy on the small Cornell light


def test_mt_and_philox_modes_agree(cornell):
    o = pyoracle.OracleScene(cornell)
    w = h = 16
    fa = np.mean([o.render("ptdirect", w * h * 2048, w, h, seed=s, rng_mode=0)[0].mean() for s in (1, 2, 3, 4)])
    fb = np.mean([o.render("ptdirect", w * h * 2048, w, h, seed=s, rng_mode=1)[0].mean() for s in (5, 6, 7, 8)])
    assert abs(fa - fb) < 0.04 * fa


def test_render_is_deterministic_in_philox_mode_and_shardable(cornell):
    o = pyoracle.OracleScene(cornell)
    w = h = 8
    n = 20000
    full, _ = o.render("ptdirect", n, w, h, seed=5, rng_mode=1, num_threads=3)
    again, _ = o.render("ptdirect", n, w, h, seed=5, rng_mode=1, num_threads=1)
    assert np.allclose(full, again, rtol=1e-12, atol=1e-15)
    a, _ = o.render("ptdirect", 12000, w, h, seed=5, rng_mode=1, sample_offset=0, film_norm_samples=n)
    b, _ = o.render("ptdirect", 8000, w, h, seed=5, rng_mode=1, sample_offset=12000, film_norm_samples=n)
    assert np.allclose(a + b, full, rtol=1e-12, atol=1e-15)


# ---- lt / ltdirect and E.area sensors (row 4): known answers ----------------------------------
def test_ltdirect_equals_ptdirect_times_light_area():
    """ltdirect estimates the same image as ptdirect EXCEPT for the reference's own quirk at src/nanogi.cpp:1017
    (`pdfPE = L->EvaluatePositionPDF(geomE)`: the LIGHT's 1/area instead of the pinhole's 1), which scales every
    contribution by the light's area. With one area light of area A: mean(ltdirect) = A * mean(ptdirect)."""
    from tests.conftest import scaled_spec
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_box(), 1.0 / 100.0), 1.0)
    area = (343 - 213) * (332 - 227) / 100.0 ** 2
    orc = pyoracle.OracleScene(sd)
    n = 48 * 48 * 256
    a, _ = orc.render("ptdirect", n, 48, 48, max_num_vertices=5, seed=1, rng_mode=1)
    b, _ = orc.render("ltdirect", n, 48, 48, max_num_vertices=5, seed=2, rng_mode=1)
    assert b.mean() / area == pytest.approx(a.mean(), rel=0.03)
    # the images agree too, not just their means (8 x 8 blocks)
    blk = lambda f: f.reshape(8, 6, 8, 6, 3).mean(axis=(1, 3, 4))
    assert np.allclose(blk(b) / area, blk(a), rtol=0.25, atol=0.02 * a.mean())


def test_light_tracing_hits_nothing_with_a_pinhole(cornell):
    """`lt` adds to the film only when a light path HITS a sensor primitive (src/nanogi.cpp:899-920); a pinhole has no
    mesh, so the reference's own `lt` image of a pinhole scene is black while all the rays are traced."""
    orc = pyoracle.OracleScene(cornell)
    f, st = orc.render("lt", 20000, 16, 16, max_num_vertices=6, seed=1, rng_mode=1)
    assert f.max() == 0.0 and st["extend_rays"] >= 20000 and st["shadow_rays"] == 0


def test_area_sensor_renderers_agree():
    """With an E.area sensor all of pt, ptdirect and lt estimate the same measurement; ltdirect is off by the factor
    InvArea(sensor) / InvArea(light) of the src/nanogi.cpp:1017 quirk."""
    from tests.conftest import scaled_spec
    sd = scenes.to_scene_data(scaled_spec(scenes.cornell_raw_sensor(), 1.0 / 100.0), 1.0)
    orc = pyoracle.OracleScene(sd)
    n = 8 * 8 * 16384
    means = {r: orc.render(r, n, 8, 8, max_num_vertices=5, seed=3 + i, rng_mode=1)[0].mean() for i, r in enumerate(["pt", "ptdirect", "lt", "ltdirect"])}
    assert means["ptdirect"] == pytest.approx(means["pt"], rel=0.03)
    assert means["lt"] == pytest.approx(means["pt"], rel=0.03)
    light_area, sensor_area = (343 - 213) * (332 - 227) / 100.0 ** 2, (200.0 / 100.0) ** 2
    assert means["ltdirect"] * sensor_area / light_area == pytest.approx(means["pt"], rel=0.04)
