import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def scene_path(name):
    return os.path.join(DATA, "scenes", name + ".xml")


@pytest.fixture(scope="session")
def cbox_desc():
    from oracle import orc
    return orc.load_scene_description(scene_path("cbox_bunny"))


@pytest.fixture(scope="session")
def native_lib():
    from psdr_cuda_b200 import build
    core = build.build_core()
    build.build_host()
    return core
