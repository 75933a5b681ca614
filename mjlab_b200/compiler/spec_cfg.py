"""Spec editors with the semantics of mjlab's ``utils/spec_config.py`` for the cfgs that
change physics: collisions (``spec_config.py:245-276``), position actuators (``:400-453``) and
contact sensors (``:514-629``).  Regex resolution follows ``utils/string.py`` (``re.match``,
first matching pattern wins, insertion order).
"""

from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any

from mjlab_b200.compiler import spec as S


def resolve_expr(pattern_map: dict[str, Any], names: list[str], default_val: Any = 0.0):
  compiled = [(re.compile(p), v) for p, v in pattern_map.items()]
  out = []
  for n in names:
    for pat, val in compiled:
      if pat.match(n):
        out.append(val)
        break
    else:
      out.append(default_val)
  return out


def filter_exp(exprs: list[str], names: list[str]) -> list[str]:
  pats = [re.compile(e) for e in exprs]
  return [n for n in names if any(p.match(n) for p in pats)]


def resolve_field(value, names, default):
  return resolve_expr(value, names, default) if isinstance(value, dict) else [value] * len(names)


_GEOM_DEFAULTS = dict(
  condim=3, contype=1, conaffinity=1, priority=0, friction=None, solref=None, solimp=None
)


@dataclass
class CollisionCfg:
  geom_names_expr: list[str]
  contype: int | dict[str, int] = 1
  conaffinity: int | dict[str, int] = 1
  condim: int | dict[str, int] = 3
  priority: int | dict[str, int] = 0
  friction: tuple | dict | None = None
  solref: tuple | dict | None = None
  solimp: tuple | dict | None = None
  disable_other_geoms: bool = True

  def validate(self) -> None:
    valid = {1, 3, 4, 6}
    vals = self.condim.values() if isinstance(self.condim, dict) else [self.condim]
    for v in vals:
      if v not in valid:
        raise ValueError(f"condim must be one of {valid}, got {v}")
    for name in ("contype", "conaffinity", "priority"):
      f = getattr(self, name)
      for v in f.values() if isinstance(f, dict) else [f]:
        if v < 0:
          raise ValueError(f"{name} must be non-negative")

  def edit_spec(self, spec: S.Spec) -> None:
    self.validate()
    all_names = [g.name for g in spec.geoms]
    subset = filter_exp(self.geom_names_expr, all_names)
    res = {k: resolve_field(getattr(self, k), subset, d) for k, d in _GEOM_DEFAULTS.items()}
    for i, name in enumerate(subset):
      g = spec.geom(name)
      g.condim, g.contype = res["condim"][i], res["contype"][i]
      g.conaffinity, g.priority = res["conaffinity"][i], res["priority"][i]
      for attr in ("friction", "solref", "solimp"):
        vals = res[attr][i]
        if vals is not None:
          arr = getattr(g, attr).copy()
          for k, v in enumerate(vals):
            arr[k] = v
          setattr(g, attr, arr)
    if self.disable_other_geoms:
      for name in set(all_names).difference(subset):
        g = spec.geom(name)
        g.contype = g.conaffinity = 0


@dataclass
class ActuatorCfg:
  joint_names_expr: list[str]
  effort_limit: float
  stiffness: float
  damping: float
  frictionloss: float = 0.0
  armature: float = 0.0


@dataclass
class ActuatorSetCfg:
  cfgs: tuple[ActuatorCfg, ...]

  def edit_spec(self, spec: S.Spec) -> None:
    for c in self.cfgs:
      if c.effort_limit <= 0:
        raise ValueError(f"effort_limit must be positive, got {c.effort_limit}")
      if c.stiffness < 0 or c.damping < 0 or c.armature < 0 or c.frictionloss < 0:
        raise ValueError("stiffness/damping/armature/frictionloss must be non-negative")
    jnts = [j for j in spec.joints if j.type != S.JNT_FREE]
    jnames = [j.name for j in jnts]
    pairs = [(c, n) for c in self.cfgs for n in filter_exp(c.joint_names_expr, jnames)]
    if self.cfgs and not pairs:
      raise ValueError("No joints matched actuator patterns")
    pairs.sort(key=lambda p: jnames.index(p[1]))
    for c, n in pairs:
      j = spec.joint(n)
      limited = j.limited == 1 or (j.limited == 2 and j.range[0] < j.range[1])
      if not limited:
        raise ValueError(f"Joint {n} must be limited for position control")
      j.armature, j.frictionloss = c.armature, c.frictionloss
      a = spec.add_actuator(
        name=n, target=n, inheritrange=1.0, forcerange=(-c.effort_limit, c.effort_limit)
      )
      a.gainprm[0] = c.stiffness
      a.biasprm[1] = -c.stiffness
      a.biasprm[2] = -c.damping


@dataclass
class ContactSensorCfg:
  name: str
  geom1: str | None = None
  body1: str | None = None
  subtree1: str | None = None
  site: str | None = None
  geom2: str | None = None
  body2: str | None = None
  subtree2: str | None = None
  num: int = 1
  data: tuple = ("found",)
  reduce: str = "none"

  def _intprm(self):
    if self.num <= 0:
      raise ValueError("'num' must be positive")
    vals = [S.CONTACT_DATA.index(k) for k in self.data] if self.data else [0]
    if any(b <= a for a, b in zip(vals, vals[1:])):
      raise ValueError(f"Data attributes must be in order: {', '.join(S.CONTACT_DATA)}")
    return (sum(1 << v for v in vals), S.CONTACT_REDUCE.index(self.reduce), self.num)

  def validate(self) -> None:
    g1 = sum(x is not None for x in (self.geom1, self.body1, self.subtree1, self.site))
    if g1 != 1:
      raise ValueError("Exactly one of geom1, body1, subtree1, or site must be specified")
    g2 = sum(x is not None for x in (self.geom2, self.body2, self.subtree2))
    if g2 > 1:
      raise ValueError("At most one of geom2, body2, subtree2 can be specified")
    if self.site is not None and g2 == 0:
      raise ValueError("Site must be used with a secondary object")

  def edit_spec(self, spec: S.Spec) -> None:
    self.validate()
    if self.geom1 is not None:
      ot, on = S.OBJ_GEOM, self.geom1
    elif self.body1 is not None:
      ot, on = S.OBJ_BODY, self.body1
    elif self.subtree1 is not None:
      ot, on = S.OBJ_XBODY, self.subtree1
    else:
      raise NotImplementedError("site-volume contact sensors are outside the hot-path subset")
    kw: dict[str, Any] = dict(name=self.name, objtype=ot, objname=on, intprm=self._intprm())
    if self.geom2 is not None:
      kw.update(reftype=S.OBJ_GEOM, refname=self.geom2)
    elif self.body2 is not None:
      kw.update(reftype=S.OBJ_BODY, refname=self.body2)
    elif self.subtree2 is not None:
      kw.update(reftype=S.OBJ_XBODY, refname=self.subtree2)
    spec.add_sensor(**kw)


@dataclass
class InitialStateCfg:
  pos: tuple = (0.0, 0.0, 0.0)
  rot: tuple = (1.0, 0.0, 0.0, 0.0)
  joint_pos: dict = field(default_factory=lambda: {".*": 0.0})
  joint_vel: dict = field(default_factory=lambda: {".*": 0.0})


@dataclass
class RobotCfg:
  """What an mjlab ``EntityCfg`` contributes to the compiled model
  (reference ``entity/entity.py:50-96,128-161``)."""

  xml: str
  init_state: InitialStateCfg
  collisions: tuple = ()
  actuators: tuple = ()
  sensors: tuple = ()
  soft_joint_pos_limit_factor: float = 1.0
  action_scale: dict = field(default_factory=dict)

  def build_spec(self) -> S.Spec:
    sp = S.Spec.from_string(self.xml)
    for c in (*self.sensors, *self.collisions):
      c.edit_spec(sp)
    if self.actuators:
      ActuatorSetCfg(tuple(self.actuators)).edit_spec(sp)
    # init_state keyframe (entity.py:146-161)
    joints = sp.joints
    comps: list[float] = []
    if joints and joints[0].type == S.JNT_FREE:
      comps += [*self.init_state.pos, *self.init_state.rot]
    non_free = [j.name for j in joints if j.type != S.JNT_FREE]
    jp = resolve_expr(self.init_state.joint_pos, non_free)
    comps += jp
    key = sp.add_key(name="init_state", qpos=comps)
    if sp.actuators:
      key.ctrl = jp
    return sp
