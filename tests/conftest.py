import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The built .so files are git-ignored: in a fresh clone compile them first (nvcc cross-compiles sm_100a without a
    GPU; about a minute).  The tests never fall back to anything else."""
    if not os.path.exists(os.path.join(ROOT, "qoaudio_b200", "libqoa_b200.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def fixture_bytes():
    with open(os.path.join(GOLDEN, "julien_baker_sprained_ankle.qoa"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)
