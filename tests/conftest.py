"""Test configuration.  `-m "not gpu"` runs on a CPU-only box (oracle vs known answers, host logic,
C-ABI symbol checks); `-m gpu` tests are the parity tests proper and go through the C-ABI."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
MODELS = os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    lib = os.path.join(ROOT, "mujoco_ros_pkgs_b200", "libb2mj.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__

        __graft_entry__.build()


_ensure_built()


@pytest.fixture(scope="session")
def capi():
    from mujoco_ros_pkgs_b200 import _capi

    return _capi


@pytest.fixture(scope="session")
def orc():
    from oracle import binding

    return binding


def model_path(name):
    return os.path.join(MODELS, name)


@pytest.fixture(scope="session")
def load_model(capi):
    cache = {}

    def _load(name):
        if name not in cache:
            cache[name] = capi.Model.from_xml_file(model_path(name))
        return cache[name]

    return _load


@pytest.fixture(scope="session")
def gpu_available(capi):
    return capi.lib.b2mj_device_count() > 0
