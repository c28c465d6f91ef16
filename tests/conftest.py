import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
# repo root (for `oracle`), the product package dir (for `gkquad_b200`) and tests/ itself
for p in (ROOT, ROOT / "gkquad-rs_b200", Path(__file__).resolve().parent):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(ROOT / "tests" / "golden" / "reference_vectors.json") as f:
        return json.load(f)
