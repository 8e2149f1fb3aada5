import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"  # present only in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer statistical test")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Host-side libraries + oracle + simulator are built once per session (no nvcc needed for CPU tests)."""
    import subprocess
    pkg = os.path.join(ROOT, "nanogi_b200")
    host_so = os.path.join(pkg, "libnanogi_host.so")
    if not os.path.exists(host_so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-o", host_so,
                               os.path.join(pkg, "host", "host_capi.cpp"), "-lz"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "liboracle.so"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostsim"), "-s", "libhostsim.so"])
    # the reference's own sources on stand-in libraries (oracle/_ref): only where /root/reference exists; tests/test_reference_pin.py
    # skips itself where it does not (the committed tests/golden/reference_films.npz still pins the oracle there)
    if os.path.exists(os.path.join(REFERENCE, "src", "nanogi.cpp")) and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libnanogi_ref.so")):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def scenes_mod():
    from nanogi_b200 import scenes
    return scenes


@pytest.fixture(scope="session")
def cornell(scenes_mod):
    return scenes_mod.to_scene_data(scenes_mod.cornell_box(), 1.0, name="cornell")


@pytest.fixture(scope="session")
def cornell_spheres(scenes_mod):
    return scenes_mod.to_scene_data(scenes_mod.cornell_spheres(), 1.0, name="cornell_spheres")


@pytest.fixture(scope="session")
def furnace(scenes_mod):
    return scenes_mod.to_scene_data(scenes_mod.furnace(0.5, 1.0), 1.0, name="furnace")


def scaled_spec(spec, scale, offset=0.0):
    """Uniformly scales/translates a scene spec (used to move scenes in and out of the fp32 'acne' regime)."""
    import copy
    out = copy.deepcopy(spec)
    for pr in out:
        if pr.get("mesh") is not None:
            pr["mesh"]["tris"] = pr["mesh"]["tris"] * scale + offset
        if "E" in pr["params"] and pr["params"]["E"]["type"] == "pinhole":
            E = pr["params"]["E"]
            E["eye"] = [x * scale + offset for x in E["eye"]]
            E["center"] = [x * scale + offset for x in E["center"]]
        if "L" in pr["params"] and pr["params"]["L"]["type"] == "point":
            pr["params"]["L"]["position"] = [x * scale + offset for x in pr["params"]["L"]["position"]]
    return out
