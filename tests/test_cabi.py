"""The C-ABI libraries load and export every symbol the headers declare (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from nanogi_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"NGI_API\s+[\w\s\*]+?\b(ngi_\w+)\s*\(", text)))


def test_host_library_exports_header_symbols():
    names = _declared("nanogi_host.h")
    assert sorted(names) == sorted(capi.HOST_SYMBOLS)
    lib = ctypes.CDLL(capi.HOST_LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_gpu_library_exports_header_symbols():
    names = _declared("nanogi_gpu.h")
    assert sorted(names) == sorted(capi.GPU_SYMBOLS)
    if not os.path.exists(capi.GPU_LIB_PATH):
        pytest.skip("libnanogi_gpu.so not built yet (needs nvcc): run __graft_entry__.build()")
    lib = ctypes.CDLL(capi.GPU_LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    lib.ngi_gpu_abi_version.restype = ctypes.c_int
    assert lib.ngi_gpu_abi_version() == capi.ABI_VERSION


def test_struct_sizes_match_the_header():
    """ctypes mirrors vs the C compiler's layout (sizes printed by a tiny C program)."""
    import subprocess, tempfile
    src = r'''
#include <stdio.h>
#include "nanogi_gpu.h"
#include "nanogi_host.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(NgiPrimitive), sizeof(NgiSceneDesc), sizeof(NgiRenderParams),
  sizeof(NgiRenderStats), sizeof(NgiSceneInfo), sizeof(NgiRay), sizeof(NgiHit), sizeof(NgiCliOptions)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    expect = [ctypes.sizeof(t) for t in (capi.NgiPrimitive, capi.NgiSceneDesc, capi.NgiRenderParams, capi.NgiRenderStats, capi.NgiSceneInfo)]
    assert sizes[:5] == expect
    assert sizes[5] == capi.RAY_DTYPE.itemsize and sizes[6] == capi.HIT_DTYPE.itemsize and sizes[7] == ctypes.sizeof(capi.NgiCliOptions)


def test_gpu_path_fails_loudly_without_a_device():
    """No CPU fallback: without a CUDA device scene_create returns NGI_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(capi.GPU_LIB_PATH):
        pytest.skip("libnanogi_gpu.so not built")
    from nanogi_b200 import scenes
    sd = scenes.to_scene_data(scenes.cornell_box(), 1.0)
    with pytest.raises(capi.NgiError, match="no CUDA device"):
        capi.GpuScene(sd, 0)
    assert capi.device_count() == -2


def test_render_flag_constants_match_the_header():
    """capi's NGI_RENDER_* mirrors vs the macros of include/nanogi_gpu.h."""
    text = open(os.path.join(ROOT, "include", "nanogi_gpu.h")).read()
    macros = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+NGI_RENDER_(\w+)\s+(\d+)u", text)}
    assert macros == {"TIME_KERNELS": capi.RENDER_TIME_KERNELS, "PER_RAY_TRACE": capi.RENDER_PER_RAY_TRACE,
                      "BDPT_PER_THREAD": capi.RENDER_BDPT_PER_THREAD}
    assert len(set(macros.values())) == len(macros) and all(v & (v - 1) == 0 for v in macros.values())      # distinct single bits


def test_multi_gpu_entry_points_fail_loudly_without_a_device():
    """ngi_gpu_group_* / ngi_gpu_comm_* (multi-GPU, NCCL): argument errors and the no-device error come back as status codes with a
    message, never as a fallback; ngi_gpu_shard_range is pure arithmetic and works anywhere."""
    import torch
    if not os.path.exists(capi.GPU_LIB_PATH):
        pytest.skip("libnanogi_gpu.so not built")
    lib = capi.gpu_lib()
    assert capi.shard_range(10, 0, 3) == (0, 3) and capi.shard_range(10, 2, 3) == (6, 4)
    h = ctypes.c_void_p()
    assert lib.ngi_gpu_group_create(None, None, 1, ctypes.byref(h)) == -1            # NGI_ERR_INVALID_ARGUMENT
    cid = capi.NgiCommId()
    assert lib.ngi_gpu_comm_create(ctypes.byref(cid), 3, 2, 0, ctypes.byref(h)) == -1      # rank >= world_size
    assert lib.ngi_gpu_comm_reduce_film(None, None, 0, 0, None) == -1
    if torch.cuda.is_available():
        return
    from nanogi_b200 import scenes
    sd = scenes.to_scene_data(scenes.cornell_box(), 1.0)
    with pytest.raises(capi.NgiError, match="no CUDA device"):
        capi.GpuGroup(sd, [0])
    rc = lib.ngi_gpu_comm_create(ctypes.byref(cid), 0, 1, 0, ctypes.byref(h))
    assert rc in (-2, -6), rc                                                          # no device (or no libnccl on this host): an error either way
