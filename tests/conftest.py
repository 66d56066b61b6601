import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "scenes"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: multi-minute CPU test")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything native is built once per session (no-op when up to date)."""
    import __graft_entry__ as entry

    entry.build()


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def golden_array(name, dtype, shape=None):
    a = np.fromfile(os.path.join(GOLDEN, name), dtype)
    return a.reshape(shape) if shape is not None else a


@pytest.fixture(scope="session")
def golden_scene():
    import rt_b200

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = rt_b200.SceneData.load(os.path.join(GOLDEN, f"{name}.rtsc"))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def scene_dir():
    """Synthetic glTF scenes generated once per session (deterministic)."""
    import gen_gltf

    d = tempfile.mkdtemp(prefix="rt_scenes_")
    made = {}

    def get(name):
        if name not in made:
            made[name] = gen_gltf.generate(name, d)
        return made[name]

    return get


@pytest.fixture(scope="session")
def big_scene(scene_dir):
    """The 260 192-triangle scene, rebuilt with the Python loader + own BVH build (bit-identical to the
    reference's flattening; test_host.py checks that on the small scenes and, when oracle/_ref exists, on this one)."""
    from rt_b200 import gltf

    sc = gltf.load_gltf(scene_dir("big_lights"), 1.0)
    sc.source_path = scene_dir("big_lights")
    return sc


def rel_mse(a, b):
    """per channel, mean over pixels of (a-b)^2 / (((a+b)/2)^2 + 1e-4) on un-tonemapped means."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return ((a - b) ** 2 / (((a + b) / 2) ** 2 + 1e-4)).reshape(-1, 3).mean(axis=0)
