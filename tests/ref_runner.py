"""Run test files of the reference (unmodified, where they lie under /root/reference/tests) against this repo's
stand-ins: ``mujoco`` / ``mujoco_warp`` / ``warp`` from mjlab_b200.compat, ``mjlab`` from baseline/_ref (tests/refload.py),
and - there being no GPU in the container - the engine compiled for the host (tests/emul/engine.py) behind the
``mujoco_warp`` stand-in.  Used by tests/test_reference_own_tests.py:  python tests/ref_runner.py <pytest args>"""
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE.parent), str(HERE), str(HERE / "emul"), str(HERE / "stubs")]  # stubs: prettytable stand-in

import mjlab_b200.compat as compat  # noqa: E402
import refload  # noqa: E402

compat.install()
sys.path.insert(0, str(refload.REF))
try:  # with the prettytable stand-in the managers package imports for real (refload would enter a namespace stub)
  import mjlab.managers  # noqa: F401
except Exception:  # noqa: BLE001
  for k in [k for k in sys.modules if k == "mjlab.managers" or k.startswith("mjlab.managers.")]:
    sys.modules.pop(k, None)
refload.load()

import mjlab_b200.compat.mujoco_warp_shim as mw  # noqa: E402
from engine import EmulEngine  # noqa: E402
from mjlab_b200.sim import native  # noqa: E402

mw._Engine = EmulEngine


def _check(rc):
  if rc:
    raise RuntimeError("b2sim call failed")


native.check = _check

import pytest  # noqa: E402

sys.exit(pytest.main(["-q", "-c", "/dev/null", "-p", "no:cacheprovider", *sys.argv[1:]]))
