"""Spec -> flat compiled :class:`Model` (the subset of ``mj_compile`` + ``mj_setConst`` the hot path needs).

Id ordering reproduces what the reference's compiler yields because ids leak into
``EntityIndexing`` (reference ``src/mjlab/entity/entity.py:588-652``): bodies depth-first in
spec order, geoms/sites/joints grouped by body in that order, dofs in joint order, actuators
in creation order.

Derived constants (``body_subtreemass``, ``dof_invweight0``, ``body_invweight0``,
``stat.meaninertia``, ``geom_rbound``, the candidate collision-pair table) are computed here in
float64 numpy at ``qpos0``; they follow MuJoCo's documented ``mj_setConst`` semantics
(restated from the public "Computation" chapter).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from mjlab_b200.compiler import spec as S
from mjlab_b200.compiler.spec import quat_mul, quat_to_mat, mat_to_quat

MINVAL = 1e-15

# (name, dtype) of every array that travels through the C ABI (include/b2sim.h: B2Array).
# "f" arrays are float64 on the host; the CUDA library narrows them to fp32 on upload.
MODEL_ARRAYS = [
  ("body_parentid", "i"), ("body_rootid", "i"), ("body_weldid", "i"),
  ("body_jntadr", "i"), ("body_jntnum", "i"), ("body_dofadr", "i"), ("body_dofnum", "i"),
  ("body_pos", "f"), ("body_quat", "f"), ("body_ipos", "f"), ("body_iquat", "f"),
  ("body_mass", "f"), ("body_subtreemass", "f"), ("body_inertia", "f"),
  ("body_invweight0", "f"),
  ("jnt_type", "i"), ("jnt_qposadr", "i"), ("jnt_dofadr", "i"), ("jnt_bodyid", "i"),
  ("jnt_limited", "i"), ("jnt_pos", "f"), ("jnt_axis", "f"), ("jnt_range", "f"),
  ("jnt_solref", "f"), ("jnt_solimp", "f"), ("jnt_margin", "f"), ("jnt_stiffness", "f"),
  ("dof_bodyid", "i"), ("dof_jntid", "i"), ("dof_parentid", "i"),
  ("dof_armature", "f"), ("dof_damping", "f"), ("dof_frictionloss", "f"),
  ("dof_invweight0", "f"),
  ("geom_type", "i"), ("geom_bodyid", "i"), ("geom_contype", "i"),
  ("geom_conaffinity", "i"), ("geom_condim", "i"), ("geom_priority", "i"),
  ("geom_size", "f"), ("geom_pos", "f"), ("geom_quat", "f"), ("geom_friction", "f"),
  ("geom_solref", "f"), ("geom_solimp", "f"), ("geom_solmix", "f"), ("geom_margin", "f"),
  ("geom_gap", "f"), ("geom_rbound", "f"), ("geom_rgba", "f"),
  ("site_bodyid", "i"), ("site_pos", "f"), ("site_quat", "f"),
  ("actuator_trnid", "i"), ("actuator_ctrllimited", "i"), ("actuator_forcelimited", "i"),
  ("actuator_gainprm", "f"), ("actuator_biasprm", "f"), ("actuator_ctrlrange", "f"),
  ("actuator_forcerange", "f"), ("actuator_gear", "f"),
  ("pair_geom1", "i"), ("pair_geom2", "i"),
  # static (world-welded, non-plane) collision geoms found through a uniform xy grid instead of the pair table
  ("static_geom", "i"), ("static_cell0", "i"), ("dyn_cgeom", "i"), ("grid_start", "i"), ("grid_items", "i"),
  ("grid_params", "f"),
  # mesh / height-field assets (csrc/b2_convex.h): geom_dataid -> mesh id (mesh geoms) or hfield id (hfield geoms)
  ("geom_dataid", "i"), ("mesh_vertadr", "i"), ("mesh_vertnum", "i"), ("mesh_vert", "f"),
  ("hfield_adr", "i"), ("hfield_nrow", "i"), ("hfield_ncol", "i"), ("hfield_size", "f"), ("hfield_data", "f"),
  ("sensor_type", "i"), ("sensor_objtype", "i"), ("sensor_objid", "i"),
  ("sensor_reftype", "i"), ("sensor_refid", "i"), ("sensor_intprm", "i"),
  ("sensor_adr", "i"), ("sensor_dim", "i"),
  ("qpos0", "f"),
]
MODEL_SCALARS_I = [
  "nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nsensor", "nsensordata", "npair", "nstatic",
  "opt_integrator", "opt_cone", "opt_solver", "opt_iterations", "opt_ls_iterations",
]
MODEL_SCALARS_F = [
  "opt_timestep", "opt_tolerance", "opt_ls_tolerance", "opt_impratio", "stat_meaninertia",
]


class _Accessor:
  def __init__(self, **kw):
    self.__dict__.update(kw)


@dataclass
class Model:
  """Compiled model: plain numpy arrays named after ``mjModel`` fields."""

  arrays: dict = field(default_factory=dict)
  names: dict = field(default_factory=dict)  # kind -> list[str]
  keys: dict = field(default_factory=dict)  # key name -> dict(qpos, qvel, ctrl)
  opt_gravity: np.ndarray = field(default_factory=lambda: np.array([0, 0, -9.81]))

  def __getattr__(self, name):
    arrays = object.__getattribute__(self, "arrays")
    if name in arrays:
      return arrays[name]
    if name in ("na", "nmocap", "ntendon", "neq"):
      return 0  # mjModel sizes of features outside the compiled subset (read by NanGuard / viewers)
    raise AttributeError(name)

  # name -> id helpers (mirror model.body(name).id style accessors, Appendix D)
  def name2id(self, kind: str, name: str) -> int:
    try:
      return self.names[kind].index(name)
    except ValueError:
      raise KeyError(f"{kind} '{name}' not found") from None

  def body(self, name):
    i = int(name) if isinstance(name, (int, np.integer)) else self.name2id("body", name)
    return _Accessor(id=i, name=self.names["body"][i], mass=np.array([float(self.body_mass[i])]),
                     pos=np.asarray(self.body_pos).reshape(-1, 3)[i], parentid=np.array([int(self.body_parentid[i])]))

  # (scalar attributes come back as 1-element arrays, as the MuJoCo bindings return them: `jnt.dofadr[0]`)
  def joint(self, name):
    i = int(name) if isinstance(name, (int, np.integer)) else self.name2id("joint", name)
    return _Accessor(
      id=i, name=self.names["joint"][i], type=np.array([int(self.jnt_type[i])]),
      qposadr=np.array([int(self.jnt_qposadr[i])]), dofadr=np.array([int(self.jnt_dofadr[i])]),
      range=np.asarray(self.jnt_range).reshape(-1, 2)[i],
    )

  def geom(self, key):
    i = int(key) if isinstance(key, (int, np.integer)) else self.name2id("geom", key)
    return _Accessor(
      id=i, name=self.names["geom"][i], condim=np.array([int(self.geom_condim[i])]),
      priority=np.array([int(self.geom_priority[i])]), friction=self.geom_friction[i],
    )

  def actuator(self, key):
    i = int(key) if isinstance(key, (int, np.integer)) else self.name2id("actuator", key)
    return _Accessor(
      id=i, name=self.names["actuator"][i], gainprm=self.actuator_gainprm[i],
      biasprm=self.actuator_biasprm[i], forcerange=self.actuator_forcerange[i],
    )

  def sensor(self, name):
    i = int(name) if isinstance(name, (int, np.integer)) else self.name2id("sensor", name)
    return _Accessor(id=i, name=self.names["sensor"][i], adr=np.array([int(self.sensor_adr[i])]),
                     dim=np.array([int(self.sensor_dim[i])]))

  def key(self, name):
    return _Accessor(**self.keys[name])

  # -- persistence (compiled blobs ship in mjlab_b200/asset_zoo/compiled) -------------------
  def save(self, path) -> None:
    import json

    meta = {
      "names": self.names,
      "keys": {k: {n: (None if v is None else list(map(float, v))) for n, v in d.items()}
               for k, d in self.keys.items()},
      "gravity": list(map(float, self.opt_gravity)),
    }
    np.savez_compressed(path, __meta__=np.array(json.dumps(meta)), **self.arrays)

  @staticmethod
  def from_binary_path(path) -> "Model":
    """``mujoco.MjModel.from_binary_path``: the blob ``mj_saveModel`` wrote (this engine's npz, whatever its suffix)."""
    return Model.load(path)

  @staticmethod
  def load(path) -> "Model":
    import json

    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["__meta__"]))
    m = Model()
    m.arrays = {k: z[k] for k in z.files if k != "__meta__"}
    fill_asset_defaults(m.arrays)
    m.names = meta["names"]
    m.keys = {
      k: {n: (None if v is None else np.array(v)) for n, v in d.items()}
      for k, d in meta["keys"].items()
    }
    m.opt_gravity = np.array(meta["gravity"])
    return m


def fill_asset_defaults(arrays: dict) -> None:
  """Model tables written before mesh / height-field assets existed (compiled blobs, MjModel-like objects
  without assets) get empty asset arrays and ``geom_dataid = -1``."""
  if "geom_dataid" not in arrays:
    arrays["geom_dataid"] = np.full(int(np.asarray(arrays["ngeom"]).reshape(-1)[0]), -1, dtype=np.int32)
  for k, dt in (("mesh_vertadr", np.int32), ("mesh_vertnum", np.int32), ("mesh_vert", np.float64),
                ("hfield_adr", np.int32), ("hfield_nrow", np.int32), ("hfield_ncol", np.int32),
                ("hfield_size", np.float64), ("hfield_data", np.float64)):
    if k not in arrays:
      arrays[k] = np.zeros((0, 3) if k == "mesh_vert" else ((0, 4) if k == "hfield_size" else 0), dtype=dt)


# ---------------------------------------------------------------------------------------
# geometry-derived inertia (bodies without <inertial>)
# ---------------------------------------------------------------------------------------


def mesh_volume_inertia(vertex: np.ndarray, face: np.ndarray | None):
  """(volume, centre of mass, 3x3 unit-density inertia about the com) of a closed triangle mesh (signed
  tetrahedra against the origin); without faces the convex hull is used, which is also what collides."""
  if face is None:
    from scipy.spatial import ConvexHull

    hull = ConvexHull(vertex)
    face = hull.simplices.copy()
    c = vertex[hull.vertices].mean(axis=0)
    for k, f in enumerate(face):  # orient outwards
      a, b, d = vertex[f]
      if np.dot(np.cross(b - a, d - a), a - c) < 0:
        face[k] = f[::-1]
  a, b, c = vertex[face[:, 0]], vertex[face[:, 1]], vertex[face[:, 2]]
  vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))
  V = vol6.sum() / 6.0
  if abs(V) < 1e-18:
    return 0.0, np.zeros(3), np.zeros((3, 3))
  com = ((a + b + c) * vol6[:, None]).sum(axis=0) / (24.0 * V)
  # second moments: integral of x x^T over a tetrahedron (0, a, b, c) = vol/20 * (sum_i p_i p_i^T + s s^T), s = a+b+c
  ssum = a + b + c
  C = (np.einsum("i,ij,ik->jk", vol6, a, a) + np.einsum("i,ij,ik->jk", vol6, b, b) + np.einsum("i,ij,ik->jk", vol6, c, c)
       + np.einsum("i,ij,ik->jk", vol6, ssum, ssum)) / 120.0
  if V < 0:
    V, C = -V, -C
  C = C - V * np.outer(com, com)
  inertia = np.trace(C) * np.eye(3) - C
  return float(V), com, inertia




def _geom_volume_inertia(g: S.Geom):
  """(volume, diagonal unit-density inertia in the geom frame)."""
  t, s = g.type, g.size
  if t == S.GEOM_SPHERE:
    r = s[0]
    v = 4.0 / 3.0 * math.pi * r**3
    return v, np.full(3, 0.4 * v * r * r)
  if t == S.GEOM_BOX:
    v = 8 * s[0] * s[1] * s[2]
    return v, v / 3.0 * np.array(
      [s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2]
    )
  if t == S.GEOM_CYLINDER:
    r, h = s[0], s[1]
    v = math.pi * r * r * 2 * h
    ixy = v * (3 * r * r + 4 * h * h) / 12.0
    return v, np.array([ixy, ixy, v * r * r / 2])
  if t == S.GEOM_ELLIPSOID:
    v = 4.0 / 3.0 * math.pi * s[0] * s[1] * s[2]
    return v, v / 5.0 * np.array(
      [s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2]
    )
  if t == S.GEOM_CAPSULE:
    r, h = s[0], s[1]
    vc = math.pi * r * r * 2 * h
    vs = 4.0 / 3.0 * math.pi * r**3
    v = vc + vs
    iz = vc * r * r / 2 + vs * 0.4 * r * r
    # hemispheres displaced by h (+ 3r/8 centroid offset inside each hemisphere)
    ixy = vc * (3 * r * r + 4 * h * h) / 12.0 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
    return v, np.array([ixy, ixy, iz])
  return 0.0, np.zeros(3)  # plane / mesh / hfield carry no mass here


def _body_inertial_from_geoms(b: S.Body, meshes: dict | None = None):
  tot_m, com = 0.0, np.zeros(3)
  parts = []
  for g in b.geoms:
    if g.type == S.GEOM_MESH and meshes and g.mesh in meshes and (g.mass is not None or g.density > 0):
      mesh = meshes[g.mesh]
      try:
        vol, mc, mi = mesh_volume_inertia(mesh.load(), mesh.face)
      except (OSError, ValueError):
        continue  # asset not readable here: the geom carries no mass (bodies of the zoo robots have <inertial>)
      if vol <= 0:
        continue
      m = g.mass if g.mass is not None else g.density * vol
      if m <= 0:
        continue
      R = quat_to_mat(g.quat)
      w, v = np.linalg.eigh(mi * (m / vol))
      parts.append((m, g.pos + R @ mc, R @ v, w))
      tot_m += m
      com += m * (g.pos + R @ mc)
      continue
    vol, inert = _geom_volume_inertia(g)
    if vol <= 0:
      continue
    m = g.mass if g.mass is not None else g.density * vol
    if m <= 0:
      continue
    parts.append((m, g.pos, quat_to_mat(g.quat), inert * (m / vol)))
    tot_m += m
    com += m * g.pos
  if tot_m <= 0:
    return np.zeros(3), np.array([1.0, 0, 0, 0]), 0.0, np.zeros(3)
  com /= tot_m
  full = np.zeros((3, 3))
  for m, p, R, inert in parts:
    d = p - com
    full += R @ np.diag(inert) @ R.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
  w, v = np.linalg.eigh(full)
  order = np.argsort(-w)  # MuJoCo sorts principal inertias in decreasing order
  w, v = w[order], v[:, order]
  if np.linalg.det(v) < 0:
    v[:, 2] *= -1
  return com, mat_to_quat(v), tot_m, w


# ---------------------------------------------------------------------------------------
# compile
# ---------------------------------------------------------------------------------------


def _rbound(t, s):
  if t == S.GEOM_SPHERE:
    return s[0]
  if t == S.GEOM_CAPSULE:
    return s[0] + s[1]
  if t == S.GEOM_CYLINDER:
    return math.hypot(s[0], s[1])
  if t in (S.GEOM_BOX, S.GEOM_ELLIPSOID):
    return float(np.linalg.norm(s)) if t == S.GEOM_BOX else float(max(s))
  return 0.0  # plane, mesh (non-colliding here)


_SUPPORTED_PAIRS = {
  (S.GEOM_PLANE, S.GEOM_SPHERE), (S.GEOM_PLANE, S.GEOM_CAPSULE), (S.GEOM_PLANE, S.GEOM_BOX),
  (S.GEOM_SPHERE, S.GEOM_SPHERE), (S.GEOM_SPHERE, S.GEOM_CAPSULE),
  (S.GEOM_CAPSULE, S.GEOM_CAPSULE),
  (S.GEOM_SPHERE, S.GEOM_BOX), (S.GEOM_CAPSULE, S.GEOM_BOX), (S.GEOM_BOX, S.GEOM_BOX),
  # convex routines (csrc/b2_convex.h): meshes through their convex hull, height fields prism by prism
  (S.GEOM_PLANE, S.GEOM_MESH), (S.GEOM_SPHERE, S.GEOM_MESH), (S.GEOM_CAPSULE, S.GEOM_MESH),
  (S.GEOM_BOX, S.GEOM_MESH), (S.GEOM_MESH, S.GEOM_MESH),
  (S.GEOM_HFIELD, S.GEOM_SPHERE), (S.GEOM_HFIELD, S.GEOM_CAPSULE), (S.GEOM_HFIELD, S.GEOM_BOX),
  (S.GEOM_HFIELD, S.GEOM_MESH),
}
# cylinders and ellipsoids: smooth convex cores with closed-form support points (GJK / EPA against every convex
# shape, prisms of a height field included; MuJoCo's plane primitives against a plane)
for _t in (S.GEOM_ELLIPSOID, S.GEOM_CYLINDER):
  _SUPPORTED_PAIRS |= {(S.GEOM_PLANE, _t), (S.GEOM_HFIELD, _t)}
  for _u in (S.GEOM_SPHERE, S.GEOM_CAPSULE, S.GEOM_ELLIPSOID, S.GEOM_CYLINDER, S.GEOM_BOX, S.GEOM_MESH):
    _SUPPORTED_PAIRS.add((min(_t, _u), max(_t, _u)))
STATIC_GRID_THRESHOLD = 16  # more static box/sphere/capsule geoms than this -> grid broadphase


def _build_static_grid(A, static, weld, gbody, gtype, ngeom, cell: float = 1.0) -> None:
  """Uniform xy grid over the static geoms' world AABBs (their bodies never move, so world pose = the
  composition of the welded chain; only world-body children are supported, which is what the terrain is).
  grid_params = [x0, y0, cell, nx, ny]; cell (ix, iy) -> index ix*ny + iy; grid_items lists positions in
  ``static_geom``; static_cell0 holds each static geom's minimal cell (ix0, iy0) for duplicate rejection."""
  A["static_geom"] = list(static)
  A["dyn_cgeom"] = []
  A["static_cell0"], A["grid_start"], A["grid_items"] = [], [0], []
  A["grid_params"] = [0.0, 0.0, cell, 0.0, 0.0]
  if not static:
    return
  lo, hi = [], []
  for g in static:
    b = gbody[g]
    if A["body_parentid"][b] != 0:
      raise NotImplementedError("static collision geoms must sit on a direct child of the world body")
    Rb = quat_to_mat(np.array(A["body_quat"][b]))
    pb = np.array(A["body_pos"][b])
    R = Rb @ quat_to_mat(np.array(A["geom_quat"][g]))
    c = pb + Rb @ np.array(A["geom_pos"][g])
    sz = np.array(A["geom_size"][g], dtype=float)
    if gtype[g] == S.GEOM_BOX:
      ext = np.abs(R) @ sz
    elif gtype[g] == S.GEOM_SPHERE:
      ext = np.full(3, sz[0])
    elif gtype[g] == S.GEOM_CAPSULE:
      ext = np.abs(R[:, 2]) * sz[1] + sz[0]
    else:
      raise NotImplementedError("static grid supports box / sphere / capsule geoms")
    lo.append(c - ext)
    hi.append(c + ext)
  lo, hi = np.array(lo), np.array(hi)
  x0, y0 = lo[:, 0].min(), lo[:, 1].min()
  # big outer borders would blow up the cell count: cap at 256 x 256 cells by growing the cell size
  span = max(hi[:, 0].max() - x0, hi[:, 1].max() - y0)
  cell = max(cell, span / 256.0)
  nx = int(np.floor((hi[:, 0].max() - x0) / cell)) + 1
  ny = int(np.floor((hi[:, 1].max() - y0) / cell)) + 1
  cells = [[] for _ in range(nx * ny)]
  for k in range(len(static)):
    ix0, ix1 = int((lo[k, 0] - x0) // cell), int((hi[k, 0] - x0) // cell)
    iy0, iy1 = int((lo[k, 1] - y0) // cell), int((hi[k, 1] - y0) // cell)
    A["static_cell0"].append([ix0, iy0])
    for ix in range(ix0, min(ix1, nx - 1) + 1):
      for iy in range(iy0, min(iy1, ny - 1) + 1):
        cells[ix * ny + iy].append(k)
  for c in cells:
    A["grid_items"].extend(c)
    A["grid_start"].append(len(A["grid_items"]))
  A["grid_params"] = [float(x0), float(y0), float(cell), float(nx), float(ny)]
  inset = set(static)
  # dynamic collision geoms that can touch the static set (bit masks are re-checked per candidate)
  smask_t = 0
  smask_a = 0
  for g in static:
    smask_t |= A["geom_contype"][g]
    smask_a |= A["geom_conaffinity"][g]
  A["dyn_cgeom"] = [g for g in range(ngeom)
                    if g not in inset and weld[gbody[g]] != 0
                    and ((A["geom_contype"][g] & smask_a) or (A["geom_conaffinity"][g] & smask_t))]
  for g in A["dyn_cgeom"]:
    if gtype[g] not in (S.GEOM_SPHERE, S.GEOM_CAPSULE, S.GEOM_ELLIPSOID, S.GEOM_CYLINDER, S.GEOM_BOX, S.GEOM_MESH):
      raise NotImplementedError("only sphere / capsule / ellipsoid / cylinder / box / mesh geoms can collide with grid-static geoms")


def compile_spec(spec: S.Spec) -> Model:
  bodies = spec.bodies
  for i, b in enumerate(bodies):
    b.id = i
  nbody = len(bodies)
  A: dict[str, list] = {k: [] for k, _ in MODEL_ARRAYS}
  names = {k: [] for k in ("body", "joint", "geom", "site", "actuator", "sensor")}

  # ---- bodies, joints, dofs ------------------------------------------------------------
  nq = nv = 0
  jid = gid = sid = 0
  dof_last_of_body = {}
  qpos0 = []
  mesh_ids: dict[str, int] = {}
  hfield_ids: dict[str, int] = {}
  for b in bodies:
    names["body"].append(b.name)
    pid = b.parent.id if b.parent is not None else 0
    A["body_parentid"].append(pid)
    A["body_pos"].append(b.pos)
    A["body_quat"].append(b.quat / np.linalg.norm(b.quat))
    if b.mass is not None:
      ipos, iquat, mass, inertia = b.ipos, b.iquat, b.mass, b.inertia
    else:
      ipos, iquat, mass, inertia = _body_inertial_from_geoms(b, spec.meshes)
    A["body_ipos"].append(ipos)
    A["body_iquat"].append(iquat / np.linalg.norm(iquat))
    A["body_mass"].append(mass)
    A["body_inertia"].append(inertia)
    A["body_jntadr"].append(jid if b.joints else -1)
    A["body_jntnum"].append(len(b.joints))
    A["body_dofadr"].append(nv if b.joints else -1)
    ndof_b = 0
    # dof chain parent: last dof of the nearest ancestor that has dofs
    anc = b.parent
    par_dof = -1
    while anc is not None:
      if anc.id in dof_last_of_body:
        par_dof = dof_last_of_body[anc.id]
        break
      anc = anc.parent
    for j in b.joints:
      j.id = jid
      names["joint"].append(j.name)
      A["jnt_type"].append(j.type)
      A["jnt_qposadr"].append(nq)
      A["jnt_dofadr"].append(nv)
      A["jnt_bodyid"].append(b.id)
      A["jnt_pos"].append(j.pos)
      A["jnt_axis"].append(j.axis)
      limited = j.limited
      if limited == 2:
        limited = int(spec.autolimits and j.range[0] < j.range[1])
      if j.type == S.JNT_FREE:
        limited = 0
      if limited and j.range[1] - j.range[0] < 2.0 * j.margin:
        # both limit rows of the joint could be active at once; the engine keeps one row per joint
        raise NotImplementedError(
          f"joint '{j.name}': range narrower than twice its margin (both limits active at once) is not supported")
      A["jnt_limited"].append(limited)
      A["jnt_range"].append(j.range)
      A["jnt_solref"].append(j.solref_limit)
      A["jnt_solimp"].append(j.solimp_limit)
      A["jnt_margin"].append(j.margin)
      A["jnt_stiffness"].append(j.stiffness)
      if j.type == S.JNT_FREE:
        n_q, n_v = 7, 6
        qpos0.extend([*b.pos, *(b.quat / np.linalg.norm(b.quat))])
      elif j.type == S.JNT_BALL:
        raise NotImplementedError("ball joints are outside the hot-path feature subset")
      else:
        n_q, n_v = 1, 1
        qpos0.append(j.ref)
      for k in range(n_v):
        A["dof_bodyid"].append(b.id)
        A["dof_jntid"].append(jid)
        A["dof_parentid"].append(par_dof)
        A["dof_armature"].append(j.armature)
        A["dof_damping"].append(j.damping)
        A["dof_frictionloss"].append(j.frictionloss)
        par_dof = nv + k
      nq += n_q
      nv += n_v
      ndof_b += n_v
      jid += 1
    if ndof_b:
      dof_last_of_body[b.id] = nv - 1
    A["body_dofnum"].append(ndof_b)
    for g in b.geoms:
      g.id = gid
      names["geom"].append(g.name)
      A["geom_type"].append(g.type)
      A["geom_bodyid"].append(b.id)
      dataid, size, rbound = -1, g.size, _rbound(g.type, g.size)
      colliding = True
      if g.type == S.GEOM_MESH:
        # A mesh collides (through its convex hull) when it has collision bits and vertex data; a mesh whose
        # file cannot be read stays what it always was here: a visual.
        colliding = False
        mesh = spec.meshes.get(g.mesh) if g.mesh else None
        if mesh is not None and (g.contype or g.conaffinity):
          try:
            v = mesh.load()
          except OSError as e:
            raise FileNotFoundError(f"geom '{g.name}': mesh '{g.mesh}' has collision bits but its file is unreadable ({e})") from None
          if g.mesh not in mesh_ids:
            mesh_ids[g.mesh] = len(mesh_ids)
            A["mesh_vertadr"].append(sum(A["mesh_vertnum"]))
            A["mesh_vertnum"].append(len(v))
            A["mesh_vert"].extend(v.tolist())
          dataid, colliding = mesh_ids[g.mesh], True
          size = np.abs(v).max(axis=0)
          rbound = float(np.linalg.norm(v, axis=1).max())
      elif g.type == S.GEOM_HFIELD:
        hf = spec.hfields.get(g.hfield) if g.hfield else None
        if hf is None:
          raise ValueError(f"geom '{g.name}': hfield asset '{g.hfield}' not found")
        if g.hfield not in hfield_ids:
          hfield_ids[g.hfield] = len(hfield_ids)
          A["hfield_adr"].append(len(A["hfield_data"]))
          A["hfield_nrow"].append(hf.nrow)
          A["hfield_ncol"].append(hf.ncol)
          A["hfield_size"].append(hf.size)
          A["hfield_data"].extend(np.asarray(hf.userdata, dtype=float).reshape(-1).tolist())
        dataid = hfield_ids[g.hfield]
        size = np.array(hf.size[:3], dtype=float)
        rbound = float(np.sqrt(hf.size[0] ** 2 + hf.size[1] ** 2 + max(hf.size[2], hf.size[3]) ** 2))
      A["geom_dataid"].append(dataid)
      A["geom_contype"].append(g.contype if colliding else 0)
      A["geom_conaffinity"].append(g.conaffinity if colliding else 0)
      A["geom_condim"].append(g.condim)
      A["geom_priority"].append(g.priority)
      A["geom_size"].append(size)
      A["geom_pos"].append(g.pos)
      A["geom_quat"].append(g.quat / np.linalg.norm(g.quat))
      A["geom_friction"].append(g.friction)
      A["geom_solref"].append(g.solref)
      A["geom_solimp"].append(g.solimp)
      A["geom_solmix"].append(g.solmix)
      A["geom_margin"].append(g.margin)
      A["geom_gap"].append(g.gap)
      A["geom_rbound"].append(rbound)
      A["geom_rgba"].append(g.rgba)
      gid += 1
    for s in b.sites:
      s.id = sid
      names["site"].append(s.name)
      A["site_bodyid"].append(b.id)
      A["site_pos"].append(s.pos)
      A["site_quat"].append(s.quat / np.linalg.norm(s.quat))
      sid += 1
  njnt, ngeom, nsite = jid, gid, sid

  parent = np.array(A["body_parentid"], dtype=np.int32)
  dofnum = np.array(A["body_dofnum"], dtype=np.int32)
  weld = np.zeros(nbody, dtype=np.int32)
  root = np.zeros(nbody, dtype=np.int32)
  for i in range(1, nbody):
    weld[i] = i if dofnum[i] > 0 else weld[parent[i]]
    root[i] = i if parent[i] == 0 else root[parent[i]]
  A["body_weldid"] = weld
  A["body_rootid"] = root

  # ---- actuators -------------------------------------------------------------------------
  for k, a in enumerate(spec.actuators):
    a.id = k
    names["actuator"].append(a.name)
    j = names["joint"].index(a.target)
    if A["jnt_type"][j] not in (S.JNT_HINGE, S.JNT_SLIDE):
      raise ValueError(f"actuator '{a.name}': only hinge/slide joint transmission is supported")
    A["actuator_trnid"].append(j)
    ctrlrange = np.array(a.ctrlrange, dtype=float)
    if a.inheritrange > 0:
      # position actuators inherit the joint range (spec_config.py:447 sets inheritrange=1.0)
      lo, hi = A["jnt_range"][j]
      mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * a.inheritrange
      ctrlrange = np.array([mid - half, mid + half])
    cl = a.ctrllimited
    if cl == 2:
      cl = int(spec.autolimits and ctrlrange[0] < ctrlrange[1])
    fl = a.forcelimited
    if fl == 2:
      fl = int(spec.autolimits and a.forcerange[0] < a.forcerange[1])
    A["actuator_ctrllimited"].append(cl)
    A["actuator_forcelimited"].append(fl)
    A["actuator_ctrlrange"].append(ctrlrange)
    A["actuator_forcerange"].append(np.array(a.forcerange, dtype=float))
    A["actuator_gainprm"].append(np.array(a.gainprm, dtype=float))
    A["actuator_biasprm"].append(np.array(a.biasprm, dtype=float))
    A["actuator_gear"].append(a.gear)
  nu = len(spec.actuators)

  # ---- sensors (contact sensors only; others keep their slot but read 0) -----------------
  adr = 0
  kind_of = {S.OBJ_BODY: "body", S.OBJ_XBODY: "body", S.OBJ_GEOM: "geom", S.OBJ_SITE: "site"}
  for k, sn in enumerate(spec.sensors):
    if sn.type == S.SENS_MJCF:  # (declared in MJCF; contact sensors arrive through add_sensor / ContactSensorCfg)
      raise NotImplementedError(
        f"sensor '{sn.name}' (<{sn.tag or 'sensor'}>): the engine evaluates contact sensors only (SURVEY.md §8a P9)")
    sn.id = k
    names["sensor"].append(sn.name)
    A["sensor_type"].append(sn.type)
    A["sensor_objtype"].append(sn.objtype)
    A["sensor_objid"].append(names[kind_of[sn.objtype]].index(sn.objname))
    A["sensor_reftype"].append(sn.reftype)
    A["sensor_refid"].append(
      names[kind_of[sn.reftype]].index(sn.refname) if sn.reftype >= 0 else -1
    )
    dataspec, reduce_, num = sn.intprm
    A["sensor_intprm"].append([dataspec, reduce_, num])
    slot = sum(d for b, d in enumerate(S.CONTACT_DATA_DIM) if dataspec >> b & 1)
    dim = slot * num
    A["sensor_adr"].append(adr)
    A["sensor_dim"].append(dim)
    adr += dim
  nsensordata = adr

  # ---- candidate collision pairs (mj_collision's static filters) -------------------------
  excl = set()
  for b1, b2 in spec.excludes:
    i1, i2 = names["body"].index(b1), names["body"].index(b2)
    excl.add((min(i1, i2), max(i1, i2)))
  gtype = A["geom_type"]
  gbody = A["geom_bodyid"]
  # static collision geoms: on a body welded to the world, not a plane, with any collision bit
  static = [g for g in range(ngeom)
            if weld[gbody[g]] == 0 and gtype[g] not in (S.GEOM_PLANE, S.GEOM_MESH, S.GEOM_HFIELD)
            and (A["geom_contype"][g] or A["geom_conaffinity"][g])]
  use_grid = len(static) > STATIC_GRID_THRESHOLD
  in_grid = set(static) if use_grid else set()
  pairs = []
  for g1 in range(ngeom):
    if g1 in in_grid:
      continue
    for g2 in range(g1 + 1, ngeom):
      if g2 in in_grid:
        continue
      ct1, ca1 = A["geom_contype"][g1], A["geom_conaffinity"][g1]
      ct2, ca2 = A["geom_contype"][g2], A["geom_conaffinity"][g2]
      if not ((ct1 & ca2) or (ct2 & ca1)):
        continue
      b1, b2 = gbody[g1], gbody[g2]
      w1, w2 = weld[b1], weld[b2]
      if w1 == w2:
        continue
      wp1, wp2 = weld[parent[w1]], weld[parent[w2]]
      if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
        continue
      if (min(b1, b2), max(b1, b2)) in excl:
        continue
      a, b = (g1, g2) if gtype[g1] <= gtype[g2] else (g2, g1)
      if (gtype[a], gtype[b]) not in _SUPPORTED_PAIRS:
        if gtype[a] in (S.GEOM_PLANE, S.GEOM_HFIELD) and gtype[b] in (S.GEOM_PLANE, S.GEOM_HFIELD):
          continue
        raise NotImplementedError(
          f"collision pair {names['geom'][a]}({gtype[a]}) - {names['geom'][b]}({gtype[b]}) "
          "is outside the primitive set of the hot path"
        )
      pairs.append((a, b))
  # contacts of one (unordered) body pair are kept contiguous: the solver accumulates one 6x6
  # block per body pair (DESIGN.md "solver").  Order inside a body pair: geom ids.
  pairs.sort(key=lambda p: (min(gbody[p[0]], gbody[p[1]]), max(gbody[p[0]], gbody[p[1]]), p))
  A["pair_geom1"] = [p[0] for p in pairs]
  A["pair_geom2"] = [p[1] for p in pairs]
  _build_static_grid(A, static if use_grid else [], weld, gbody, gtype, ngeom)

  # features the engine does not implement must fail here, not silently change the physics
  used = {g for p in pairs for g in p}
  for g in used:
    if A["geom_condim"][g] not in (1, 3):
      raise NotImplementedError(
        f"geom '{names['geom'][g]}': condim {A['geom_condim'][g]} (torsional/rolling friction) is not implemented")
  if any(f != 0 for f in A["dof_frictionloss"]):
    raise NotImplementedError("joint frictionloss (dry-friction constraint rows) is not implemented")
  A["qpos0"] = qpos0
  m = Model()
  shapes = {
    "body_pos": 3, "body_quat": 4, "body_ipos": 3, "body_iquat": 4, "body_inertia": 3,
    "jnt_pos": 3, "jnt_axis": 3, "jnt_range": 2, "jnt_solref": 2, "jnt_solimp": 5,
    "geom_size": 3, "geom_pos": 3, "geom_quat": 4, "geom_friction": 3, "geom_solref": 2,
    "geom_solimp": 5, "geom_rgba": 4, "site_pos": 3, "site_quat": 4,
    "actuator_gainprm": 10, "actuator_biasprm": 10, "actuator_ctrlrange": 2,
    "actuator_forcerange": 2, "sensor_intprm": 3, "body_invweight0": 2, "static_cell0": 2,
    "mesh_vert": 3, "hfield_size": 4,
  }
  for k, dt in MODEL_ARRAYS:
    if k in ("body_subtreemass", "body_invweight0", "dof_invweight0"):
      continue
    arr = np.array(A[k], dtype=np.float64 if dt == "f" else np.int32)
    if k in shapes:
      arr = arr.reshape(-1, shapes[k])
    m.arrays[k] = np.ascontiguousarray(arr)
  o = spec.option
  m.opt_gravity = np.array(o.gravity, dtype=float)
  scal = dict(
    nq=nq, nv=nv, nu=nu, nbody=nbody, njnt=njnt, ngeom=ngeom, nsite=nsite,
    nsensor=len(spec.sensors), nsensordata=nsensordata, npair=len(pairs), nstatic=len(A["static_geom"]),
    opt_integrator=o.integrator, opt_cone=o.cone, opt_solver=o.solver,
    opt_iterations=o.iterations, opt_ls_iterations=o.ls_iterations,
    opt_timestep=o.timestep, opt_tolerance=o.tolerance, opt_ls_tolerance=o.ls_tolerance,
    opt_impratio=o.impratio,
  )
  for k, v in scal.items():
    m.arrays[k] = np.array(v, dtype=np.int32 if k in MODEL_SCALARS_I else np.float64)
  m.names = names
  if nv > 64 or nbody > 64:
    raise NotImplementedError("hot path supports nv <= 64 and nbody <= 64 (dof/body bitmasks)")
  _set_const(m)
  for k in spec.keys:
    qp = k.qpos if k.qpos is not None else m.qpos0
    qv, ct = k.qvel, k.ctrl
    if k.scope is not None and len(qp) != nq:
      # a key of an attached child: its values sit at the child's joints / actuators, defaults elsewhere
      jn, an = names["joint"], names["actuator"]
      full_q, full_v, full_c = np.array(m.qpos0, dtype=float), np.zeros(nv), np.zeros(nu)
      width = {0: (7, 6), 1: (4, 3), 2: (1, 1), 3: (1, 1)}
      iq = iv = 0
      for name in k.scope[0]:
        j = jn.index(name)
        wq, wv = width[int(A["jnt_type"][j])]
        qa, va = int(A["jnt_qposadr"][j]), int(A["jnt_dofadr"][j])
        if iq + wq <= len(qp):
          full_q[qa:qa + wq] = np.asarray(qp, dtype=float)[iq:iq + wq]
        if qv is not None and iv + wv <= len(qv):
          full_v[va:va + wv] = np.asarray(qv, dtype=float)[iv:iv + wv]
        iq, iv = iq + wq, iv + wv
      if ct is not None:
        for i, name in enumerate(k.scope[1]):
          if i < len(ct):
            full_c[an.index(name)] = float(ct[i])
      qp, qv, ct = full_q, full_v, full_c
    if len(qp) != nq:
      raise ValueError(f"key '{k.name}': qpos has {len(qp)} entries, model nq={nq}")
    m.keys[k.name] = dict(
      qpos=np.array(qp, dtype=float),
      qvel=np.zeros(nv) if qv is None else np.array(qv, dtype=float),
      ctrl=np.zeros(nu) if ct is None or len(ct) != nu else np.array(ct, float),
    )
  return m


# ---------------------------------------------------------------------------------------
# mj_setConst subset: quantities evaluated at qpos0
# ---------------------------------------------------------------------------------------


def kinematics_qpos0(m: Model, qpos=None):
  """Numpy forward kinematics (used only for compile-time constants and host utilities)."""
  nb = int(m.nbody)
  qpos = m.qpos0 if qpos is None else qpos
  xpos = np.zeros((nb, 3))
  xquat = np.zeros((nb, 4))
  xquat[0, 0] = 1
  xanchor = np.zeros((int(m.njnt), 3))
  xaxis = np.zeros((int(m.njnt), 3))
  for i in range(1, nb):
    p = m.body_parentid[i]
    jn, ja = m.body_jntnum[i], m.body_jntadr[i]
    if jn == 1 and m.jnt_type[ja] == S.JNT_FREE:
      a = m.jnt_qposadr[ja]
      pos = qpos[a : a + 3].copy()
      quat = qpos[a + 3 : a + 7] / np.linalg.norm(qpos[a + 3 : a + 7])
      xanchor[ja] = pos
      xaxis[ja] = quat_to_mat(quat) @ m.jnt_axis[ja]
    else:
      R = quat_to_mat(xquat[p])
      pos = xpos[p] + R @ m.body_pos[i]
      quat = quat_mul(xquat[p], m.body_quat[i])
      for j in range(ja, ja + jn):
        Rq = quat_to_mat(quat)
        xanchor[j] = pos + Rq @ m.jnt_pos[j]
        xaxis[j] = Rq @ m.jnt_axis[j]
        d = qpos[m.jnt_qposadr[j]] - m.qpos0[m.jnt_qposadr[j]]
        if m.jnt_type[j] == S.JNT_SLIDE:
          pos = pos + xaxis[j] * d
        else:
          ql = np.concatenate([[math.cos(d / 2)], m.jnt_axis[j] * math.sin(d / 2)])
          quat = quat_mul(quat, ql)
          pos = xanchor[j] - quat_to_mat(quat) @ m.jnt_pos[j]
    xpos[i] = pos
    xquat[i] = quat / np.linalg.norm(quat)
  return xpos, xquat, xanchor, xaxis


def mass_matrix_qpos0(m: Model):
  """Dense joint-space inertia at qpos0 plus what the invweight computation needs."""
  nb, nv = int(m.nbody), int(m.nv)
  xpos, xquat, xanchor, xaxis = kinematics_qpos0(m)
  xipos = np.zeros((nb, 3))
  ximat = np.zeros((nb, 3, 3))
  for i in range(nb):
    R = quat_to_mat(xquat[i])
    xipos[i] = xpos[i] + R @ m.body_ipos[i]
    ximat[i] = quat_to_mat(quat_mul(xquat[i], m.body_iquat[i]))
  # 6D motion vectors [ang; lin] of each dof expressed at the world origin
  Sdof = np.zeros((nv, 6))
  for j in range(int(m.njnt)):
    d, b = m.jnt_dofadr[j], m.jnt_bodyid[j]
    t = m.jnt_type[j]
    if t == S.JNT_FREE:
      R = quat_to_mat(xquat[b])
      for k in range(3):
        Sdof[d + k, 3 + k] = 1.0
        ax = R[:, k]
        Sdof[d + 3 + k, :3] = ax
        Sdof[d + 3 + k, 3:] = np.cross(ax, -xpos[b])
    elif t == S.JNT_HINGE:
      Sdof[d, :3] = xaxis[j]
      Sdof[d, 3:] = np.cross(xaxis[j], -xanchor[j])
    else:
      Sdof[d, 3:] = xaxis[j]
  # chain membership
  chain = np.zeros((nb, nv), dtype=bool)
  for i in range(1, nb):
    chain[i] = chain[m.body_parentid[i]]
    if m.body_dofnum[i] > 0:
      chain[i, m.body_dofadr[i] : m.body_dofadr[i] + m.body_dofnum[i]] = True
  M = np.diag(np.array(m.dof_armature, dtype=float)) if nv else np.zeros((0, 0))
  for i in range(1, nb):
    mass = m.body_mass[i]
    if mass <= 0:
      continue
    Ic = ximat[i] @ np.diag(m.body_inertia[i]) @ ximat[i].T
    c = xipos[i]
    cx = np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])
    I6 = np.zeros((6, 6))
    I6[:3, :3] = Ic + mass * (cx @ cx.T)
    I6[:3, 3:] = mass * cx
    I6[3:, :3] = mass * cx.T
    I6[3:, 3:] = mass * np.eye(3)
    J = Sdof * chain[i][:, None]
    M += J @ I6 @ J.T
  return M, Sdof, chain, xipos


def _set_const(m: Model) -> None:
  nb, nv = int(m.nbody), int(m.nv)
  sub = np.array(m.body_mass, dtype=float).copy()
  for i in range(nb - 1, 0, -1):
    sub[m.body_parentid[i]] += sub[i]
  m.arrays["body_subtreemass"] = sub
  inv_b = np.zeros((nb, 2))
  inv_d = np.zeros(nv)
  mean = 1.0
  if nv:
    M, Sdof, chain, xipos = mass_matrix_qpos0(m)
    Minv = np.linalg.inv(M)
    mean = float(np.mean(np.diag(M)))
    for i in range(1, nb):
      if not chain[i].any():
        continue
      c = xipos[i]
      # 6 x nv Jacobian of body i at its com: rows [lin(3); ang(3)] in MuJoCo's jacp/jacr order
      J = np.zeros((6, nv))
      for d in np.nonzero(chain[i])[0]:
        ang, lin0 = Sdof[d, :3], Sdof[d, 3:]
        J[:3, d] = lin0 + np.cross(ang, c)
        J[3:, d] = ang
      Ainv = J @ Minv @ J.T
      inv_b[i, 0] = max(MINVAL, np.trace(Ainv[:3, :3]) / 3.0)
      inv_b[i, 1] = max(MINVAL, np.trace(Ainv[3:, 3:]) / 3.0)
    dinv = np.diag(Minv).copy()
    for j in range(int(m.njnt)):
      d = m.jnt_dofadr[j]
      if m.jnt_type[j] == S.JNT_FREE:
        dinv[d : d + 3] = dinv[d : d + 3].mean()
        dinv[d + 3 : d + 6] = dinv[d + 3 : d + 6].mean()
    inv_d = dinv
  m.arrays["body_invweight0"] = inv_b
  m.arrays["dof_invweight0"] = inv_d
  m.arrays["stat_meaninertia"] = np.array(mean, dtype=np.float64)
