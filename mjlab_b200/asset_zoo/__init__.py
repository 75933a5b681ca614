"""Benchmark robots (Unitree G1 / Go1) and scene assembly.

The MJCF files themselves are the reference's data (``src/mjlab/asset_zoo/robots/*/xmls``) and
are *not* copied into this repository.  ``tools/compile_assets.py`` compiles them (in the
authoring container, where ``/root/reference`` exists) into the flat model blobs under
``compiled/`` which is what tests, ``bench.py`` and ``smoke()`` load on the GPU box.
"""

from __future__ import annotations

import os
from pathlib import Path

from mjlab_b200.compiler.compile import Model

COMPILED_DIR = Path(__file__).parent / "compiled"
_REF_ZOO = Path(
  os.environ.get("MJLAB_ASSET_ZOO", "/root/reference/src/mjlab/asset_zoo/robots")
)


def reference_xml(robot: str) -> str:
  """MJCF text of a zoo robot, read from the reference checkout (authoring container only)."""
  p = _REF_ZOO / f"unitree_{robot}" / "xmls" / f"{robot}.xml"
  if not p.exists():
    raise FileNotFoundError(
      f"{p} not found: the MJCF sources live in the reference checkout; on a GPU box use "
      "load_compiled() (blobs produced by tools/compile_assets.py)"
    )
  return p.read_text()


def load_compiled(name: str) -> Model:
  """Load a pre-compiled scene model, e.g. ``g1_flat``, ``g1_tracking_flat``, ``go1_flat``."""
  p = COMPILED_DIR / f"{name}.npz"
  if not p.exists():
    raise FileNotFoundError(f"{p} missing: run `python tools/compile_assets.py`")
  return Model.load(p)
