import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


GPU_TEST_TIMEOUT_S = 900  # the slowest GPU test (full-size configs[2] properties) takes about a minute on a B200


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a kernel that never ends, a lost device) must end the run instead of holding the box:
    every `gpu` test without its own limit gets one.  method="thread": a blocked CUDA call never returns to the
    interpreter, so a signal handler would never run; the watchdog thread dumps the stacks and exits the process."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(GPU_TEST_TIMEOUT_S, method="thread"))


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
