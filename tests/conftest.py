import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# every device array of the library gets guard bands while the tests run (checked when it is freed and at the end of the
# session; compute-sanitizer is not available on the GPU pool).  Tools the tests start inherit the setting.
os.environ.setdefault("MGIC_ARENA_GUARD", "4096")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import mg_ic_code_b200 as m
    c = m.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session", autouse=True)
def _guard_bands():
    """no kernel of the session wrote past either end of an array"""
    yield
    if "mg_ic_code_b200" not in sys.modules:
        return
    import mg_ic_code_b200 as m
    violations, _ = m.Context.guard_check()
    assert violations == 0, f"{violations} device arrays had their guard bands overwritten (see stderr)"
