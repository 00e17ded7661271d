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
