import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stm32-speech-recognition_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build (or re-use) every native artefact once per session; cheap when up to date."""
    import __graft_entry__ as g
    if os.environ.get("SR_NO_BUILD") == "1":     # builder's GPU runs: use the .so files that travelled with the snapshot
        return
    g.build()


@pytest.fixture(scope="session")
def handle():
    import sr_b200
    h = sr_b200.Handle(0)
    yield h
    h.close()
