import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_usable():
    """A CUDA device AND the built library: without either, the gpu-marked tests are skipped (the CPU suite
    — oracle KATs, goldens, host logic, gloo — still runs); on a GPU box a missing library is an error."""
    try:
        from amclj_b200 import _lib
        c = _lib.Context(0)
        c.close()
        return True, ""
    except Exception as e:  # AmcljError(ERR_NODEVICE), ImportError (library not built), OSError
        return False, str(e)


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    ok, why = _gpu_usable()
    if ok:
        return
    import shutil
    if shutil.which("nvidia-smi"):  # a GPU box: never skip silently there
        return
    skip = pytest.mark.skip(reason=f"no usable CUDA device here ({why})")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "amclj_golden.json").read_text())


@pytest.fixture(scope="session")
def lgfloor():
    from amclj_b200 import maps
    return maps.lgfloor()


@pytest.fixture(scope="session")
def fixture_scan(golden):
    s = golden["scan"]
    return dict(ranges=np.asarray(s["ranges"], np.float32), angle_min=float(s["angleMin"]),
                angle_inc=float(s["angleIncrement"]), range_max=float(s["rangeMax"]))
