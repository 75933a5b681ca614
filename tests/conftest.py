import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p_ in (str(ROOT), str(ROOT / "tests")):
  if p_ not in sys.path:
    sys.path.insert(0, p_)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def g1_model():
  from mjlab_b200.asset_zoo import load_compiled

  return load_compiled("g1_flat")


@pytest.fixture(scope="session")
def go1_model():
  from mjlab_b200.asset_zoo import load_compiled

  return load_compiled("go1_flat")
