import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def hb():
    """The product package (ctypes over the two in-tree .so files)."""
    return importlib.import_module("hai719-raytracing_b200")


@pytest.fixture(scope="session")
def ref():
    """The oracle = the reference itself, built headless (oracle/_ref). Skips if it is not built."""
    import oracle_ref
    if not oracle_ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle_ref.Ref()


@pytest.fixture(scope="session")
def assets():
    d = os.path.join(ROOT, "assets", "_ref")
    if not os.path.isdir(os.path.join(d, "mesh")):
        pytest.skip("assets/_ref not staged")
    return d


GOLDEN = os.path.join(ROOT, "tests", "golden")
ALL_SCENES = ["single_sphere", "single_square", "cornell_box", "mesh", "rt_in_a_weekend", "random_spheres",
              "debug_refraction", "flamingo", "raccoon", "flamingo_pond", "backrooms_pool", "flamingo_lake", "config5"]
