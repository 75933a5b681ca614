"""Stand-ins for the three modules mjlab's Python layer imports around the hot path — ``mujoco_warp``,
``warp`` and ``mujoco`` — backed by libb2sim.so, so that the reference's *own* ``mjlab.sim.Simulation``
(``src/mjlab/sim/sim.py:94-198``), ``WarpBridge``/``TorchArray`` (``sim/sim_data.py``), ``expand_model_fields``
(``sim/randomization.py``), ``EntityData``/``Entity.initialize`` (``entity/data.py``, ``entity/entity.py:325-423``)
and event functions (``envs/mdp/events.py``) run unmodified on the B200 engine.

  import mjlab_b200.compat as compat
  compat.install()            # registers the stand-ins in sys.modules (only for modules that are absent)
  from mjlab.sim import Simulation, SimulationCfg     # the reference's class, now stepping libb2sim

Scope: exactly the surface those files touch at the engine boundary —
``mjwarp.put_model / put_data / step / forward``, ``Model``/``Data`` structs whose array fields are ``wp.array``
views of engine memory, the ``wp`` device/stream/graph/array plumbing they call, and the ``mujoco`` enums,
``MjModel`` accessors and ``MjData``/``mj_forward`` used around construction.  Rendering, ``MjSpec`` editing of
meshes/textures/cameras and the Warp kernel language are out of scope (``wp.launch`` only understands
element-wise kernels written against ``wp.tid()``, which is what ``repeat_array_kernel`` is).
"""

from __future__ import annotations

import importlib
import importlib.util
import sys


def _present(name: str) -> bool:
  if name in sys.modules:
    return not getattr(sys.modules[name], "__b2_compat__", False)
  try:
    return importlib.util.find_spec(name) is not None
  except (ImportError, ValueError):
    return False


def install(force: bool = False) -> list[str]:
  """Register the stand-ins for whichever of warp / mujoco_warp / mujoco cannot be imported.
  Returns the names that were shimmed."""
  done = []
  for public, private in (("warp", "warp_shim"), ("mujoco", "mujoco_shim"), ("mujoco_warp", "mujoco_warp_shim")):
    if not force and _present(public):
      continue
    mod = importlib.import_module(f"mjlab_b200.compat.{private}")
    sys.modules[public] = mod
    done.append(public)
  return done
