"""Randomised scenes through the host emulation of the CUDA sources against the fp64 oracle (tools/emul_fuzz.py):
synthetic floating-base robots (up to 64 dofs, hinge and slide joints, limits, margins / gaps, condim 1, solref /
solimp variants, force-limited actuators, both integrators, self-collisions) and free primitives among static
obstacles (pair table + static grid + every primitive pair type).  A seed may hit a configuration that no
implementation defines (two capsule axes that intersect exactly: zero-length contact normal; a capsule along a box face
at the length threshold of the one-or-two-contacts rule) - those are reported by the tool and bounded here."""

import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(*args):
  r = subprocess.run([sys.executable, str(ROOT / "tools" / "emul_fuzz.py"), "--keep-going", *args],
                     capture_output=True, text=True, timeout=900)
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
  return r.stdout


def test_fuzzed_robots_match_the_oracle():
  out = _run("--seeds", "60", "--worlds", "2", "--steps", "2", "--first", "1000")
  last = out.strip().splitlines()[-1]
  bad = re.findall(r"\((\d+), ", last) if last.startswith("FAILED") else []
  assert len(bad) <= 2, last          # degenerate geometry (see the docstring), never a trend
  errs = [float(x) for x in re.findall(r"max rel err ([0-9.e+-]+)", out)]
  assert len(errs) >= 58 and sorted(errs)[len(errs) // 2] < 1e-4, (len(errs), sorted(errs)[-5:])


def test_fuzzed_obstacle_scenes_match_the_oracle():
  out = _run("--scene", "obstacle", "--seeds", "40")
  last = out.strip().splitlines()[-1]
  ties = re.findall(r"\d+", last) if last.startswith("contact-count ties") else []
  assert len(ties) <= 2, last
  errs = [float(x) for x in re.findall(r"worst qacc rel err ([0-9.e+-]+)", out)]
  assert len(errs) >= 38 and max(errs) < 1e-2, (len(errs), sorted(errs)[-5:])
