import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SAMPLE_TXT = os.path.join(GOLDEN, "sample.txt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def sd():
    """The product package (loads libsyldet_cuda.so; builds it first if the tree is fresh)."""
    lib = os.path.join(ROOT, "syllable-detector-swift_b200", "libsyldet_cuda.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("syldet_b200")


@pytest.fixture(scope="session")
def cw():
    return importlib.import_module("syllable-detector-swift_b200.config_writer")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("tools.synth")


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(GOLDEN, "golden_cases.npz"))
    cases = {}
    for name in z["names"]:
        name = str(name)
        cases[name] = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + ".")}
        cases[name]["config"] = bytes(cases[name]["config"]).decode()
    return cases


@pytest.fixture(scope="session")
def sample_text():
    return open(SAMPLE_TXT).read()
