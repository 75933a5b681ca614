"""Scene assembly: terrain + robot -> one compiled model.

Mirrors ``Scene.__init__`` of the reference (``src/mjlab/scene/scene.py:24-40,133-147``): the
terrain spec is attached first and un-prefixed, each entity with the prefix ``"<name>/"``, so the
flat-terrain G1 scene has body 0 ``world``, 1 ``terrain``, 2.. ``robot/pelvis`` subtree and geom 0
is the ground plane (``terrains/terrain_importer.py:154-163``).  The per-env origin sites that
``terrain_importer.py:98-120`` adds are *not* materialised (SURVEY.md §7 "hard parts": they are
static world-body sites and would cost 800 MB at 4096 envs); env origins stay a host tensor.
Box terrains (``terrains.RoughTerrainCfg``) become grid-static geoms of the body ``terrain``.
"""

from __future__ import annotations

from mjlab_b200.compiler import spec as S
from mjlab_b200.compiler.compile import Model
from mjlab_b200.compiler.spec_cfg import RobotCfg


def flat_terrain_spec(name: str = "terrain") -> S.Spec:
  sp = S.Spec()
  sp.worldbody.add_body(name=name).add_geom(
    name=name, type=S.GEOM_PLANE, size=(0, 0, 0.01)
  )
  return sp


def build_scene_spec(robot: RobotCfg, terrain="plane", robot_name: str = "robot"):
  """Returns (spec, terrain_origins or None).  ``terrain`` is ``"plane"``, ``None`` or a
  ``terrains.RoughTerrainCfg`` (box terrain: one static body ``terrain`` holding every box,
  ``terrain_importer.py:125-152``)."""
  from mjlab_b200.terrains import RoughTerrainCfg, terrain_spec

  sp = S.Spec()
  origins = None
  if terrain == "plane":
    sp.attach(flat_terrain_spec(), prefix="")
  elif isinstance(terrain, RoughTerrainCfg):
    tsp, origins = terrain_spec(terrain)
    sp.attach(tsp, prefix="")
  elif terrain is not None:
    raise NotImplementedError(f"terrain '{terrain}' (plane and box terrains are on the hot path)")
  sp.attach(robot.build_spec(), prefix=f"{robot_name}/")
  return sp, origins


def apply_mujoco_cfg(sp: S.Spec, cfg) -> None:
  """``MujocoCfg.edit_spec`` (reference ``sim/sim.py:65-82``)."""
  o = sp.option
  o.timestep, o.impratio = cfg.timestep, cfg.impratio
  o.iterations, o.tolerance = cfg.iterations, cfg.tolerance
  o.ls_iterations, o.ls_tolerance = cfg.ls_iterations, cfg.ls_tolerance
  o.gravity = tuple(cfg.gravity)
  o.integrator = {"euler": S.INT_EULER, "implicitfast": S.INT_IMPLICITFAST}[cfg.integrator]
  o.cone = {"pyramidal": S.CONE_PYRAMIDAL, "elliptic": S.CONE_ELLIPTIC}[cfg.cone]
  o.solver = {"newton": S.SOL_NEWTON, "cg": S.SOL_CG, "pgs": S.SOL_PGS}[cfg.solver]


def compile_scene(robot: RobotCfg, mujoco_cfg=None, terrain="plane") -> Model:
  sp, origins = build_scene_spec(robot, terrain)
  if mujoco_cfg is not None:
    apply_mujoco_cfg(sp, mujoco_cfg)
  m = sp.compile()
  if origins is not None:
    m.arrays["terrain_origins"] = origins  # (num_rows, num_cols, 3) spawn points, host-side use only
  return m
