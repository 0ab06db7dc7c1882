import json
import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return json.loads((GOLDEN / name).read_text())


def _gpu_available():
    try:
        from phamclust_b200 import _native
        ctx = _native.Context(0)
        ctx.close()
        return True, ""
    except Exception as exc:          # no GPU / no library
        return False, str(exc)


@pytest.fixture(scope="session")
def gpu_ctx():
    """One pc_ctx for the whole GPU session.  A missing library or GPU is a
    FAILURE under -m gpu, never a silent skip to some other path."""
    from phamclust_b200 import _native
    ctx = _native.Context(int(os.environ.get("PC_TEST_DEVICE", "0")))
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def reference():
    from oracle import reference_loader
    if not reference_loader.available():
        pytest.skip("/root/reference is not mounted here (expected on the GPU box)")
    return reference_loader.load()
