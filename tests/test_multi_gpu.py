"""The product's multi-GPU entry points (include/nanogi_gpu.h: ngi_gpu_group_*, ngi_gpu_comm_*, ngi_gpu_shard_range): samples
sharded by index, scene built once and broadcast, ONE NCCL reduce of the per-GPU films (reference gather: src/nanogi.cpp:429-437).
The 2-GPU cases need `gpurun --gpus 2`; on a 1-GPU box they skip and the single-device group still runs."""
import ctypes

import numpy as np
import pytest

from nanogi_b200 import capi, scenes, shard


def test_shard_range_is_the_products_arithmetic():
    """ngi_gpu_shard_range (no device needed): contiguous, exhaustive, sizes within one of each other."""
    for n in (0, 1, 7, 1000, 2123366400, 8493465600, 2**62 + 12345):
        for g in (1, 2, 3, 8):
            parts = [capi.shard_range(n, r, g) for r in range(g)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (o0, c0), (o1, _) in zip(parts, parts[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
            assert parts == [shard.shard_range(n, r, g) for r in range(g)]


@pytest.mark.gpu
@pytest.mark.parametrize("renderer", ["pt", "ptdirect"])
def test_group_of_one_device_equals_the_scene_handle(renderer):
    sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
    n, w = 1 << 20, 64
    g = capi.GpuGroup(sd, [0])
    fg, sg = g.render(renderer, n, w, w, seed=21)
    s = capi.GpuScene(sd, 0)
    fs, ss = s.render(renderer, n, w, w, seed=21)
    assert sg.extend_rays == ss.extend_rays and sg.shadow_rays == ss.shadow_rays and sg.paths == n
    assert np.allclose(fg, fs, rtol=2e-4, atol=1e-6 * fs.max())          # same samples; fp32 atomics in another order
    g.close(); s.close()


@pytest.mark.gpu
def test_film_is_accumulated_in_fp64_behind_the_fp32_atomics():
    """A 1 x 1 image: every splat of 2^26 ptdirect samples (~1.3e8 adds of ~1e-8 of the final value each) lands on ONE pixel.
    A plain fp32 sum stops growing once a splat is below ulp(sum)/2 (6e-8 of the sum) and comes out far too dark; the module
    folds its fp32 staging film into a fp64 accumulator every wavefront iteration (k_film_fold), like the reference's double
    film (src/nanogi.cpp:203, :297), so the pixel must agree with the CPU oracle's fp64 estimate."""
    from oracle import pyoracle
    sd = scenes.to_scene_data(scenes.cornell_box(), 1.0)
    s = capi.GpuScene(sd, 0)
    big, _ = s.render("ptdirect", 1 << 26, 1, 1, max_num_vertices=6, seed=5)
    small, _ = s.render("ptdirect", 1 << 20, 1, 1, max_num_vertices=6, seed=6)          # few enough adds for plain fp32
    ref, _ = pyoracle.OracleScene(sd).render("ptdirect", 1 << 20, 1, 1, max_num_vertices=6, seed=7)
    # (a plain fp32 sum of 1.3e8 such terms stalls around 1/3 of the true value; the tolerances below are Monte-Carlo noise of
    # the 2^20-sample runs)
    assert abs(small.mean() - ref.mean()) < 0.03 * ref.mean()
    assert abs(big.mean() - small.mean()) < 0.03 * small.mean() and abs(big.mean() - ref.mean()) < 0.03 * ref.mean(), (big.mean(), small.mean(), ref.mean())
    s.close()


@pytest.mark.gpu
def test_accumulate_adds_passes_into_the_callers_film():
    """ngi_gpu_render_device with accumulate = 1: eight calls over eighths of the sample range equal one call over the range."""
    import torch
    sd = scenes.to_scene_data(scenes.cornell_box(), 1.0)
    s = capi.GpuScene(sd, 0)
    n, w = 1 << 22, 64
    one, _ = s.render("ptdirect", n, w, w, seed=5)
    film = torch.zeros((w, w, 3), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()                                   # stream 0 below = the module's own (non-blocking) stream
    for k in range(8):
        s.render_device(film.data_ptr(), 0, "ptdirect", n // 8, w, w, seed=5, sample_offset=k * (n // 8), film_norm_samples=n, accumulate=1)
    torch.cuda.synchronize()
    assert np.allclose(film.cpu().numpy(), one, rtol=1e-4, atol=1e-6 * one.max())
    s.close()


@pytest.mark.gpu
def test_two_gpu_group_equals_one_gpu():
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    sd = scenes.to_scene_data(scenes.cornell_spheres(), 1.0)
    n, w = 1 << 22, 96
    g2 = capi.GpuGroup(sd, [0, 1])
    f2, s2 = g2.render("ptdirect", n, w, w, seed=33)
    g1 = capi.GpuGroup(sd, [1])                    # the second device alone: its scene came over ncclBroadcast in g2, here it is built locally
    f1, s1 = g1.render("ptdirect", n, w, w, seed=33)
    assert s2.extend_rays == s1.extend_rays and s2.shadow_rays == s1.shadow_rays      # the same sample SET whatever the GPU count
    assert s2.reduce_seconds > 0
    assert np.allclose(f2, f1, rtol=2e-4, atol=1e-6 * f1.max())
    # closest hits through the broadcast copy of the BVH are bit-identical to the locally built one
    h = ctypes.c_void_p()
    assert g2.lib.ngi_gpu_group_scene(g2.handle, 1, ctypes.byref(h)) == 0
    rays = np.concatenate([scenes.camera_rays(sd, 64, 64), scenes.random_rays(sd, 1 << 14, 5)])
    hits_b = np.empty(rays.shape[0], capi.HIT_DTYPE)
    capi._check(g2.lib.ngi_gpu_trace(h, rays.ctypes.data, rays.shape[0], hits_b.ctypes.data, 0, 0), "ngi_gpu_trace")
    s0 = capi.GpuScene(sd, 0)
    hits_a = s0.trace(rays)
    assert np.array_equal(hits_a, hits_b)
    g2.close(); g1.close(); s0.close()
