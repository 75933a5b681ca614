"""``mujoco`` stand-in for the part of the module mjlab's entity / sim layer touches around the hot path
(SURVEY.md App. D): the enums (values are MuJoCo's public constants), ``MjModel`` (the compiled model of
``mjlab_b200.compiler``: ``mjModel`` field names as numpy attributes, ``model.joint(name).dofadr`` style
accessors), ``MjSpec`` (the MJCF-subset spec of ``mjlab_b200.compiler.spec``), ``MjData`` and ``mj_forward`` as the
host-side placeholders ``Simulation.__init__`` creates (``sim/sim.py:105-107``; the engine runs its own forward
at creation), and the three state functions ``NanGuard`` calls when enabled (``utils/nan_guard.py:58-77,129``).
"""

from __future__ import annotations

import enum

import numpy as np

from mjlab_b200.compiler.compile import Model as MjModel
from mjlab_b200.compiler.spec import Actuator as MjsActuator
from mjlab_b200.compiler.spec import Body as MjsBody
from mjlab_b200.compiler.spec import Geom as MjsGeom
from mjlab_b200.compiler.spec import Joint as MjsJoint
from mjlab_b200.compiler.spec import Site as MjsSite
from mjlab_b200.compiler.spec import Spec as MjSpec

__b2_compat__ = True
__version__ = "3.3.7-b2compat"


class mjtJoint(enum.IntEnum):
  mjJNT_FREE = 0
  mjJNT_BALL = 1
  mjJNT_SLIDE = 2
  mjJNT_HINGE = 3


class mjtGeom(enum.IntEnum):
  mjGEOM_PLANE = 0
  mjGEOM_HFIELD = 1
  mjGEOM_SPHERE = 2
  mjGEOM_CAPSULE = 3
  mjGEOM_ELLIPSOID = 4
  mjGEOM_CYLINDER = 5
  mjGEOM_BOX = 6
  mjGEOM_MESH = 7
  mjGEOM_SDF = 8


class mjtObj(enum.IntEnum):
  mjOBJ_UNKNOWN = 0
  mjOBJ_BODY = 1
  mjOBJ_XBODY = 2
  mjOBJ_JOINT = 3
  mjOBJ_DOF = 4
  mjOBJ_GEOM = 5
  mjOBJ_SITE = 6
  mjOBJ_CAMERA = 7
  mjOBJ_LIGHT = 8
  mjOBJ_ACTUATOR = 19
  mjOBJ_SENSOR = 20


class mjtSensor(enum.IntEnum):
  mjSENS_TOUCH = 0
  mjSENS_ACCELEROMETER = 1
  mjSENS_VELOCIMETER = 2
  mjSENS_GYRO = 3
  mjSENS_FORCE = 4
  mjSENS_TORQUE = 5
  mjSENS_MAGNETOMETER = 6
  mjSENS_RANGEFINDER = 7
  mjSENS_CAMPROJECTION = 8
  mjSENS_JOINTPOS = 9
  mjSENS_JOINTVEL = 10
  mjSENS_TENDONPOS = 11
  mjSENS_TENDONVEL = 12
  mjSENS_ACTUATORPOS = 13
  mjSENS_ACTUATORVEL = 14
  mjSENS_ACTUATORFRC = 15
  mjSENS_JOINTACTFRC = 16
  mjSENS_TENDONACTFRC = 17
  mjSENS_BALLQUAT = 18
  mjSENS_BALLANGVEL = 19
  mjSENS_JOINTLIMITPOS = 20
  mjSENS_JOINTLIMITVEL = 21
  mjSENS_JOINTLIMITFRC = 22
  mjSENS_TENDONLIMITPOS = 23
  mjSENS_TENDONLIMITVEL = 24
  mjSENS_TENDONLIMITFRC = 25
  mjSENS_FRAMEPOS = 26
  mjSENS_FRAMEQUAT = 27
  mjSENS_FRAMEXAXIS = 28
  mjSENS_FRAMEYAXIS = 29
  mjSENS_FRAMEZAXIS = 30
  mjSENS_FRAMELINVEL = 31
  mjSENS_FRAMEANGVEL = 32
  mjSENS_FRAMELINACC = 33
  mjSENS_FRAMEANGACC = 34
  mjSENS_SUBTREECOM = 35
  mjSENS_SUBTREELINVEL = 36
  mjSENS_SUBTREEANGMOM = 37
  mjSENS_INSIDESITE = 38
  mjSENS_GEOMDIST = 39
  mjSENS_GEOMNORMAL = 40
  mjSENS_GEOMFROMTO = 41
  mjSENS_CONTACT = 42
  mjSENS_E_POTENTIAL = 43
  mjSENS_E_KINETIC = 44
  mjSENS_CLOCK = 45
  mjSENS_TACTILE = 46
  mjSENS_PLUGIN = 47
  mjSENS_USER = 48


class mjtGain(enum.IntEnum):
  mjGAIN_FIXED = 0
  mjGAIN_AFFINE = 1
  mjGAIN_MUSCLE = 2
  mjGAIN_USER = 3


class mjtBias(enum.IntEnum):
  mjBIAS_NONE = 0
  mjBIAS_AFFINE = 1
  mjBIAS_MUSCLE = 2
  mjBIAS_USER = 3


class mjtDyn(enum.IntEnum):
  mjDYN_NONE = 0
  mjDYN_INTEGRATOR = 1
  mjDYN_FILTER = 2
  mjDYN_FILTEREXACT = 3
  mjDYN_MUSCLE = 4
  mjDYN_USER = 5


class mjtTrn(enum.IntEnum):
  mjTRN_JOINT = 0
  mjTRN_JOINTINPARENT = 1
  mjTRN_SLIDERCRANK = 2
  mjTRN_TENDON = 3
  mjTRN_SITE = 4
  mjTRN_BODY = 5


class mjtIntegrator(enum.IntEnum):
  mjINT_EULER = 0
  mjINT_RK4 = 1
  mjINT_IMPLICIT = 2
  mjINT_IMPLICITFAST = 3


class mjtCone(enum.IntEnum):
  mjCONE_PYRAMIDAL = 0
  mjCONE_ELLIPTIC = 1


class mjtJacobian(enum.IntEnum):
  mjJAC_DENSE = 0
  mjJAC_SPARSE = 1
  mjJAC_AUTO = 2


class mjtSolver(enum.IntEnum):
  mjSOL_PGS = 0
  mjSOL_CG = 1
  mjSOL_NEWTON = 2


class mjtLimited(enum.IntEnum):
  mjLIMITED_FALSE = 0
  mjLIMITED_TRUE = 1
  mjLIMITED_AUTO = 2


class mjtState(enum.IntFlag):
  mjSTATE_TIME = 1
  mjSTATE_QPOS = 2
  mjSTATE_QVEL = 4
  mjSTATE_ACT = 8
  mjSTATE_PHYSICS = 14
  mjSTATE_FULLPHYSICS = 15


class MjData:
  """Host-side single-world state (what ``Simulation.mj_data`` hands to viewers): qpos at ``qpos0``."""

  def __init__(self, model: MjModel):
    self.qpos = np.array(model.qpos0, dtype=np.float64)
    self.qvel = np.zeros(int(model.nv))
    self.act = np.zeros(0)
    self.ctrl = np.zeros(int(model.nu))
    self.time = 0.0


def mj_forward(model: MjModel, data: MjData) -> None:
  """No host-side pipeline exists here: derived quantities are produced by the engine (``b2_create`` runs a
  forward pass over every world), so the host placeholder is left at the state it was given."""
  return None


def mj_resetData(model: MjModel, data: MjData) -> None:
  """Host-side ``MjData`` back to the model's defaults (``qpos0``, zero velocity and control)."""
  data.qpos = np.array(model.qpos0, dtype=np.float64)
  data.qvel = np.zeros(int(model.nv))
  data.ctrl = np.zeros(int(model.nu))
  data.time = 0.0


def mj_resetDataKeyframe(model: MjModel, data: MjData, key: int) -> None:
  """State of keyframe ``key`` (by position in the model's keyframe list) into the host-side ``MjData``."""
  k = list(model.keys.values())[key]
  data.qpos = np.array(k["qpos"] if k.get("qpos") is not None else model.qpos0, dtype=np.float64)
  data.qvel = np.array(k["qvel"], dtype=np.float64) if k.get("qvel") is not None else np.zeros(int(model.nv))
  if k.get("ctrl") is not None:
    data.ctrl = np.array(k["ctrl"], dtype=np.float64)
  data.time = 0.0


def mj_stateSize(model: MjModel, spec: int) -> int:
  n = 0
  if spec & mjtState.mjSTATE_TIME: n += 1
  if spec & mjtState.mjSTATE_QPOS: n += int(model.nq)
  if spec & mjtState.mjSTATE_QVEL: n += int(model.nv)
  return n


def mj_getState(model: MjModel, data, state: np.ndarray, spec: int) -> None:
  parts = []
  if spec & mjtState.mjSTATE_TIME: parts.append(np.atleast_1d(float(data.time)))
  if spec & mjtState.mjSTATE_QPOS: parts.append(np.asarray(data.qpos, dtype=np.float64).ravel())
  if spec & mjtState.mjSTATE_QVEL: parts.append(np.asarray(data.qvel, dtype=np.float64).ravel())
  state[:] = np.concatenate(parts) if parts else np.zeros(0)


def mj_saveModel(model: MjModel, filename: str, buffer=None) -> None:
  """The compiled-model blob of this engine (``Model.save``: npz) written under exactly the requested name
  (``utils/nan_guard.py`` stores ``model_<stamp>.mjb`` next to its dump and looks it up by that name)."""
  with open(filename, "wb") as f:
    model.save(f)


# visual enums referenced at import time by utils/spec_config.py (texture / light / camera editors); the
# engine ignores visual elements, the values are MuJoCo's
class mjtTexture(enum.IntEnum):
  mjTEXTURE_2D = 0
  mjTEXTURE_CUBE = 1
  mjTEXTURE_SKYBOX = 2


class mjtBuiltin(enum.IntEnum):
  mjBUILTIN_NONE = 0
  mjBUILTIN_GRADIENT = 1
  mjBUILTIN_CHECKER = 2
  mjBUILTIN_FLAT = 3


class mjtMark(enum.IntEnum):
  mjMARK_NONE = 0
  mjMARK_EDGE = 1
  mjMARK_CROSS = 2
  mjMARK_RANDOM = 3


class mjtCamLight(enum.IntEnum):
  mjCAMLIGHT_FIXED = 0
  mjCAMLIGHT_TRACK = 1
  mjCAMLIGHT_TRACKCOM = 2
  mjCAMLIGHT_TARGETBODY = 3
  mjCAMLIGHT_TARGETBODYCOM = 4


class mjtLightType(enum.IntEnum):
  mjLIGHT_SPOT = 0
  mjLIGHT_DIRECTIONAL = 1
  mjLIGHT_POINT = 2
  mjLIGHT_IMAGE = 3


class mjtTextureRole(enum.IntEnum):
  mjTEXROLE_USER = 0
  mjTEXROLE_RGB = 1
  mjTEXROLE_OCCLUSION = 2
  mjTEXROLE_ROUGHNESS = 3
  mjTEXROLE_METALLIC = 4
  mjTEXROLE_NORMAL = 5
  mjTEXROLE_OPACITY = 6
  mjTEXROLE_EMISSIVE = 7
  mjTEXROLE_RGBA = 8
  mjTEXROLE_ORM = 9


class _SpecElement:
  """One element of an attached entity spec after compilation: the name it carries in the scene (with the
  entity prefix) and its global id, which is what ``Entity._compute_indexing`` reads (``entity.py:588-601``)."""

  def __init__(self, name: str, id: int, **kw):
    self.name, self.id = name, id
    self.__dict__.update(kw)

  def __repr__(self):
    return f"<{self.name}:{self.id}>"


class EntitySpecView:
  """What ``Entity.spec`` looks like once ``Scene`` has attached it under ``prefix`` and compiled the scene
  (``scene/scene.py:133-147``): element lists restricted to the entity, ids global.  Built from a compiled
  model, for callers that hold only the compiled blob (the MJCF sources are not shipped to the GPU box)."""

  def __init__(self, model: MjModel, prefix: str = "robot/"):
    def pick(kind, **extra):
      out = []
      for i, n in enumerate(model.names.get(kind, [])):
        if n.startswith(prefix):
          out.append(_SpecElement(n, i, **{k: v(i) for k, v in extra.items()}))
      return out

    self.prefix = prefix
    # bodies[0] plays the entity's own worldbody (skipped by `spec.bodies[1:]`)
    self.bodies = [_SpecElement(prefix + "world", 0)] + pick("body")
    self.joints = pick("joint", type=lambda i: mjtJoint(int(model.jnt_type[i])))
    body_in = [n.startswith(prefix) for n in model.names["body"]]

    def on_entity(kind, owner):  # unnamed elements belong to the entity that owns their body
      names = model.names.get(kind, [])
      return [_SpecElement(names[i] if i < len(names) else "", i) for i in range(len(owner)) if body_in[int(owner[i])]]

    self.geoms = on_entity("geom", model.geom_bodyid)
    self.sites = on_entity("site", model.site_bodyid)
    self.sensors = pick("sensor")
    self.actuators = pick("actuator")
    self.tendons = []
    self.keys = []

  def add_key(self, **kw):
    k = _SpecElement(self.prefix + kw.get("name", "key"), len(self.keys), **{a: b for a, b in kw.items() if a != "name"})
    self.keys.append(k)
    return k
