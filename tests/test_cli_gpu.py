"""The drop-in `nanogi` command line (nanogi_b200/host/nanogi_main.cpp) end to end on the GPU box: reference CLI in,
film file out — Run / Renderer::Load / Renderer::Render / RenderProcess's pass loop (reference src/nanogi.cpp:1993-2121,
:117-221, :225-440)."""
import os
import subprocess

import numpy as np
import pytest

from nanogi_b200 import capi, scenes

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nanogi_b200", "nanogi")


def run(*args):
    r = subprocess.run([BIN, *[str(a) for a in args]], capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout + r.stderr


@pytest.fixture(scope="module")
def scene_file(tmp_path_factory):
    d = tmp_path_factory.mktemp("cornell")
    return scenes.write_scene_files(scenes.cornell_box(), str(d))


@pytest.mark.parametrize("renderer", ["pt", "ptdirect", "ltdirect", "bdpt"])
def test_cli_renders_like_the_c_abi(scene_file, tmp_path, renderer):
    out = tmp_path / f"{renderer}.pfm"
    rc, log = run(renderer, scene_file, out, 64, 64, "-n", 64 * 64 * 64, "-m", 6, "--seed", 7)
    assert rc == 0, log
    img = capi.load_image(str(out))[::-1]                      # file is top-down, film is bottom-up
    sd = capi.load_scene_file(scene_file, 1.0)
    g = capi.GpuScene(sd, 0)
    film, _ = g.render(renderer, 64 * 64 * 64, 64, 64, max_num_vertices=6, seed=7)
    g.close()
    assert np.allclose(img, film, rtol=1e-4, atol=1e-6 * film.max())


def test_cli_multi_gpu_shards_sum_to_the_single_gpu_film(scene_file, tmp_path):
    """--gpus 2 (one host thread per device, samples sharded by index, films added on the host): the same sample set as one GPU."""
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    n = 1 << 24
    one, two = tmp_path / "one.pfm", tmp_path / "two.pfm"
    assert run("ptdirect", scene_file, one, 64, 64, "-n", n, "-m", 8, "--seed", 13)[0] == 0
    rc, log = run("ptdirect", scene_file, two, 64, 64, "-n", n, "-m", 8, "--seed", 13, "--gpus", 2)
    assert rc == 0, log
    a, b = capi.load_image(str(one)), capi.load_image(str(two))
    assert np.allclose(a, b, rtol=2e-3, atol=1e-5 * a.max())


def test_cli_unsupported_renderers_fail_loudly(scene_file, tmp_path):
    rc, log = run("ptmnee", scene_file, tmp_path / "x.hdr", 16, 16, "-n", 100)
    assert rc != 0 and "not supported" in log
    rc, log = run("nonsense", scene_file, tmp_path / "x.hdr", 16, 16)
    assert rc != 0 and "Invalid renderer type" in log


def test_cli_render_time_and_progress_images(scene_file, tmp_path):
    """--render-time runs passes until the time is up (src/nanogi.cpp:336-346) and progress images are written between
    passes with the {{count}} template (:356-404)."""
    out = tmp_path / "t.hdr"
    fmt = str(tmp_path / "progress" / "{{count}}.hdr")
    rc, log = run("ptdirect", scene_file, out, 128, 128, "-t", 1.5, "--progress-image-update-interval", 0.3,
                  "--progress-image-update-format", fmt, "--seed", 3)
    assert rc == 0, log
    assert out.exists()
    assert (tmp_path / "progress").is_dir(), log
    shots = sorted(os.listdir(tmp_path / "progress"))
    assert len(shots) >= 2 and shots[0] == "0000000001.hdr", log
    n = [int(line.split(":")[-1]) for line in log.splitlines() if "# of samples" in line][0]
    assert n > 10_000_000                                       # far more than a CPU would do in 1.5 s
    a = capi.load_image(str(tmp_path / "progress" / shots[0])); b = capi.load_image(str(out))
    assert abs(a.mean() - b.mean()) < 0.1 * b.mean()            # progress images are normalised by the samples so far


def test_cli_resume_continues_the_sample_sequence(scene_file, tmp_path):
    """--sample-offset / --resume-from: two runs of N samples equal one run of 2N (counter-based RNG: the sample set is
    the same; only the fp32 summation order differs)."""
    n = 1 << 22
    a, b, c = tmp_path / "a.pfm", tmp_path / "b.pfm", tmp_path / "c.pfm"
    assert run("ptdirect", scene_file, a, 64, 64, "-n", n, "--seed", 9)[0] == 0
    rc, log = run("ptdirect", scene_file, b, 64, 64, "-n", n, "--seed", 9, "--sample-offset", n, "--resume-from", a)
    assert rc == 0, log
    assert run("ptdirect", scene_file, c, 64, 64, "-n", 2 * n, "--seed", 9)[0] == 0
    fb, fc = capi.load_image(str(b)), capi.load_image(str(c))
    assert np.allclose(fb, fc, rtol=2e-3, atol=1e-5 * fc.max())
    fa = capi.load_image(str(a))
    assert not np.allclose(fa, fc, rtol=2e-3, atol=1e-5 * fc.max())


def test_cli_matches_the_reference_binary(scene_file, tmp_path):
    """Same command line into the reference application itself (oracle/_ref/nanogi_ref: src/nanogi.cpp's own main on stand-in
    libraries, CPU) and into the drop-in (GPU): the two images agree within Monte-Carlo noise, block by block."""
    from oracle import pyref
    if not os.path.exists(pyref.BIN_PATH):
        pytest.skip("oracle/_ref/nanogi_ref not built")
    n, w = 64 * 64 * 256, 64
    ref_img, gpu_img = tmp_path / "ref.exr", tmp_path / "gpu.pfm"
    r = subprocess.run([pyref.BIN_PATH, "ptdirect", scene_file, str(ref_img), str(w), str(w), "-n", str(n), "-m", "6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and ref_img.exists(), r.stdout[-2000:]
    rc, log = run("ptdirect", scene_file, gpu_img, w, w, "-n", n, "-m", 6, "--seed", 11)
    assert rc == 0, log
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    a = cv2.imread(str(ref_img), cv2.IMREAD_UNCHANGED)[:, :, ::-1].astype(np.float64)     # BGR -> RGB, top-down
    b = capi.load_image(str(gpu_img)).astype(np.float64)
    assert a.shape == b.shape == (w, w, 3)
    assert abs(a.mean() - b.mean()) < 0.02 * a.mean()
    blk = lambda f: f.reshape(8, 8, 8, 8, 3).mean(axis=(1, 3, 4))
    assert np.allclose(blk(a), blk(b), rtol=0.15, atol=0.02 * a.mean())


def test_reference_binary_with_the_gpu_bridge(scene_file, tmp_path):
    """INTEGRATION.md B, compiled: oracle/_ref/nanogi_ref_gpu is the REFERENCE application (its Run, CLI, Scene::Load, SaveImage) with
    oracle/ref_gpu_bridge.hpp's RenderOnGpu behind Renderer::Render (one inserted statement, src/nanogi.cpp:203), linked to
    libnanogi_gpu.so. NANOGI_DEVICE=gpu selects the GPU module; without it the same binary runs the reference's CPU path. The GPU image
    equals the drop-in `nanogi`'s (same seed: same Philox sample set) and agrees with the CPU image within Monte-Carlo noise."""
    from oracle import pyref
    exe = os.path.join(os.path.dirname(pyref.BIN_PATH), "nanogi_ref_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/nanogi_ref_gpu not built")
    n, w = 64 * 64 * 256, 64
    cpu_img, gpu_img, drop_img = tmp_path / "cpu.exr", tmp_path / "gpu.exr", tmp_path / "drop.pfm"
    args = ["ptdirect", scene_file, None, str(w), str(w), "-n", str(n), "-m", "6"]
    r = subprocess.run([exe] + [a if a is not None else str(cpu_img) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and cpu_img.exists(), r.stdout[-2000:]
    env = dict(os.environ, NANOGI_DEVICE="gpu", NANOGI_GPUS="1", NANOGI_SEED="11")
    r = subprocess.run([exe] + [a if a is not None else str(gpu_img) for a in args], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and gpu_img.exists() and "GPU module:" in r.stdout + r.stderr, (r.stdout + r.stderr)[-2000:]
    rc, log = run("ptdirect", scene_file, drop_img, w, w, "-n", n, "-m", 6, "--seed", 11)
    assert rc == 0, log
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    a = cv2.imread(str(cpu_img), cv2.IMREAD_UNCHANGED)[:, :, ::-1].astype(np.float64)     # BGR -> RGB, top-down
    b = cv2.imread(str(gpu_img), cv2.IMREAD_UNCHANGED)[:, :, ::-1].astype(np.float64)
    c = capi.load_image(str(drop_img)).astype(np.float64)
    assert a.shape == b.shape == c.shape == (w, w, 3)
    # the reference's loader (Assimp stand-in) and this repository's loader hand the module the same scene: same samples, same film
    assert np.allclose(b, c, rtol=2e-3, atol=1e-5 * c.max())
    assert abs(a.mean() - b.mean()) < 0.02 * a.mean()
    blk = lambda f: f.reshape(8, 8, 8, 8, 3).mean(axis=(1, 3, 4))
    assert np.allclose(blk(a), blk(b), rtol=0.15, atol=0.02 * a.mean())
