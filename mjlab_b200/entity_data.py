"""`EntityData` / `EntityIndexing` — the accessor layer through which everything above the physics
boundary reads and writes simulation state.

Restates the reference's ``src/mjlab/entity/data.py:20-516`` and ``entity/entity.py:19-47,588-652`` over
our bridges: same property names, shapes and frame conventions —

* quaternions ``wxyz``; ``xfrc_applied[..., 0:3]`` force, ``[..., 3:6]`` torque (``data.py:150-153``);
* ``cvel = [ang(0:3), lin(3:6)]`` with the linear part taken at ``subtree_com[root]``, converted to the
  velocity of a point ``pos`` by ``lin - ang x (subtree_com - pos)`` (``data.py:20-31``);
* free-joint ``qvel``: linear velocity in the world frame, angular velocity in the body frame
  (``envs/mdp/events.py:87,142``).

Indices are ``torch.int32`` and env ids are reshaped to ``(N, 1)`` for advanced indexing exactly as the
reference does (``data.py:180-188``); both work on the engine's strided views.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import torch

from mjlab_b200.compiler import spec as S
from mjlab_b200.compiler.compile import Model


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
  aw, ax, ay, az = a.unbind(-1)
  bw, bx, by, bz = b.unbind(-1)
  return torch.stack(
    [aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], dim=-1)


def quat_apply(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
  w, u = q[..., 0:1], q[..., 1:4]
  t = 2.0 * torch.cross(u, v, dim=-1)
  return v + w * t + torch.cross(u, t, dim=-1)


def quat_apply_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
  w, u = q[..., 0:1], q[..., 1:4]
  t = 2.0 * torch.cross(u, v, dim=-1)
  return v - w * t + torch.cross(u, t, dim=-1)


def quat_from_matrix(m: torch.Tensor) -> torch.Tensor:
  """Rotation matrices ``(..., 3, 3)`` -> unit quaternions ``(..., 4)`` (w >= 0 branch-free form)."""
  m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
  qa = torch.stack(
    [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1
  ).clamp(min=0.0).sqrt()
  cand = torch.stack(
    [
      torch.stack([qa[..., 0] ** 2, m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1]], -1),
      torch.stack([m[..., 2, 1] - m[..., 1, 2], qa[..., 1] ** 2, m[..., 1, 0] + m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0]], -1),
      torch.stack([m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] + m[..., 0, 1], qa[..., 2] ** 2, m[..., 1, 2] + m[..., 2, 1]], -1),
      torch.stack([m[..., 1, 0] - m[..., 0, 1], m[..., 2, 0] + m[..., 0, 2], m[..., 2, 1] + m[..., 1, 2], qa[..., 3] ** 2], -1),
    ], dim=-2)
  cand = cand / (2.0 * qa[..., None].clamp(min=0.1))
  best = qa.argmax(dim=-1)
  q = torch.gather(cand, -2, best[..., None, None].expand(*best.shape, 1, 4)).squeeze(-2)
  q = torch.where(q[..., 0:1] < 0, -q, q)
  return q / q.norm(dim=-1, keepdim=True)


def compute_velocity_from_cvel(pos: torch.Tensor, subtree_com: torch.Tensor, cvel: torch.Tensor) -> torch.Tensor:
  """World-frame ``[lin, ang]`` velocity of the point ``pos`` from a com-based ``cvel`` (``data.py:20-31``)."""
  lin_c, ang = cvel[..., 3:6], cvel[..., 0:3]
  lin_w = lin_c - torch.cross(ang, subtree_com - pos, dim=-1)
  return torch.cat([lin_w, ang], dim=-1)


@dataclass(frozen=True)
class EntityIndexing:
  """Global ids / addresses of one entity's elements (``entity/entity.py:19-47``)."""

  body_names: tuple
  joint_names: tuple
  geom_names: tuple
  site_names: tuple
  actuator_names: tuple
  body_ids: torch.Tensor
  geom_ids: torch.Tensor
  site_ids: torch.Tensor
  ctrl_ids: torch.Tensor
  joint_ids: torch.Tensor
  joint_q_adr: torch.Tensor
  joint_v_adr: torch.Tensor
  free_joint_q_adr: torch.Tensor
  free_joint_v_adr: torch.Tensor
  sensor_adr: dict = field(default_factory=dict)
  root_body_id: int = 0

  @staticmethod
  def from_model(model: Model, entity: str = "robot", device: str = "cpu") -> "EntityIndexing":
    """Elements whose compiled name carries the prefix ``"<entity>/"`` (``scene/scene.py:133-147``), in id
    order, with the prefix stripped from the reported names (``entity/entity.py:189-214``)."""
    pre = f"{entity}/"

    def pick(kind):
      return [(i, n[len(pre):]) for i, n in enumerate(model.names[kind]) if n.startswith(pre)]

    def t(v):
      return torch.tensor(v, dtype=torch.int, device=device)

    bodies, acts = pick("body"), pick("actuator")
    own = {i for i, _ in bodies}

    def on_entity(kind, owner):  # unnamed elements (e.g. visual mesh geoms) carry no prefix: go by body
      return [(i, n[len(pre):] if n.startswith(pre) else n) for i, n in enumerate(model.names[kind])
              if int(owner[i]) in own]

    geoms, sites = on_entity("geom", model.geom_bodyid), on_entity("site", model.site_bodyid)
    joints = pick("joint")
    jq, jv, fq, fv, jid, jn = [], [], [], [], [], []
    for j, name in joints:
      qa, va = int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
      if int(model.jnt_type[j]) == S.JNT_FREE:
        fq += range(qa, qa + 7)
        fv += range(va, va + 6)
      else:
        jq.append(qa)
        jv.append(va)
        jid.append(j)
        jn.append(name)
    sens = {}
    for i, n in enumerate(model.names["sensor"]):
      if n.startswith(pre):
        a, d = int(model.sensor_adr[i]), int(model.sensor_dim[i])
        sens[n[len(pre):]] = torch.arange(a, a + d, dtype=torch.int, device=device)
    if not bodies:
      raise ValueError(f"no bodies with prefix '{pre}' in the model")
    return EntityIndexing(
      body_names=tuple(n for _, n in bodies), joint_names=tuple(jn),
      geom_names=tuple(n for _, n in geoms), site_names=tuple(n for _, n in sites),
      actuator_names=tuple(n for _, n in acts),
      body_ids=t([i for i, _ in bodies]), geom_ids=t([i for i, _ in geoms]),
      site_ids=t([i for i, _ in sites]), ctrl_ids=t([i for i, _ in acts]), joint_ids=t(jid),
      joint_q_adr=t(jq), joint_v_adr=t(jv), free_joint_q_adr=t(fq), free_joint_v_adr=t(fv),
      sensor_adr=sens, root_body_id=bodies[0][0],
    )


class EntityData:
  """State accessors of one entity over ``sim.data`` / ``sim.model`` (``entity/data.py:34-516``)."""

  ROOT_POSE_DIM, ROOT_VEL_DIM, ROOT_STATE_DIM = 7, 6, 13

  def __init__(self, indexing: EntityIndexing, data, model, device: str, num_envs: int,
               default_root_state: torch.Tensor | None = None,
               default_joint_pos: torch.Tensor | None = None,
               soft_joint_pos_limit_factor: float = 1.0, native: bool = True):
    self.indexing, self.data, self.model, self.device = indexing, data, model, device
    self.native = native
    ix = indexing
    self.is_fixed_base = ix.free_joint_q_adr.numel() == 0
    self.is_articulated = ix.joint_q_adr.numel() > 0
    self.is_actuated = ix.ctrl_ids.numel() > 0
    nj = ix.joint_q_adr.numel()
    f32 = dict(dtype=torch.float32, device=device)
    self.default_root_state = (
      default_root_state if default_root_state is not None
      else torch.tensor([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], **f32).repeat(num_envs, 1))
    self.default_joint_pos = default_joint_pos if default_joint_pos is not None else torch.zeros(num_envs, nj, **f32)
    self.default_joint_vel = torch.zeros(num_envs, nj, **f32)
    self.gravity_vec_w = torch.tensor([0.0, 0.0, -1.0], **f32).repeat(num_envs, 1)
    self.forward_vec_b = torch.tensor([1.0, 0.0, 0.0], **f32).repeat(num_envs, 1)
    if nj:
      lim = model.jnt_range[:, ix.joint_ids]  # (num_envs, nj, 2)  (entity.py:372)
      self.joint_pos_limits = lim.clone() if isinstance(lim, torch.Tensor) else torch.as_tensor(lim).clone()
      mid = self.joint_pos_limits.mean(dim=-1)
      half = 0.5 * (self.joint_pos_limits[..., 1] - self.joint_pos_limits[..., 0]) * soft_joint_pos_limit_factor
      self.soft_joint_pos_limits = torch.stack([mid - half, mid + half], dim=-1)
      self.default_joint_pos_limits = self.joint_pos_limits.clone()

  # -- writers (data.py:69-178) ---------------------------------------------------------------------
  def _resolve_env_ids(self, env_ids):
    # data.py:180-188: None -> all envs; tensor -> column for broadcasting; slices pass through unchanged
    if env_ids is None:
      return slice(None)
    if isinstance(env_ids, torch.Tensor):
      return env_ids[:, None]
    return env_ids

  def write_root_state(self, root_state, env_ids=None) -> None:
    if self.is_fixed_base:
      raise ValueError("Cannot write root state for fixed-base entity.")
    assert root_state.shape[-1] == self.ROOT_STATE_DIM
    self.write_root_pose(root_state[:, :7], env_ids)
    self.write_root_velocity(root_state[:, 7:], env_ids)

  def write_root_pose(self, pose, env_ids=None) -> None:
    if self.is_fixed_base:
      raise ValueError("Cannot write root pose for fixed-base entity.")
    assert pose.shape[-1] == self.ROOT_POSE_DIM
    self.data.qpos[self._resolve_env_ids(env_ids), self.indexing.free_joint_q_adr] = pose

  def write_root_velocity(self, velocity, env_ids=None) -> None:
    if self.is_fixed_base:
      raise ValueError("Cannot write root velocity for fixed-base entity.")
    assert velocity.shape[-1] == self.ROOT_VEL_DIM
    self.data.qvel[self._resolve_env_ids(env_ids), self.indexing.free_joint_v_adr] = velocity

  def write_joint_state(self, position, velocity, joint_ids=None, env_ids=None) -> None:
    if not self.is_articulated:
      raise ValueError("Cannot write joint state for non-articulated entity.")
    self.write_joint_position(position, joint_ids, env_ids)
    self.write_joint_velocity(velocity, joint_ids, env_ids)

  def write_joint_position(self, position, joint_ids=None, env_ids=None) -> None:
    if not self.is_articulated:
      raise ValueError("Cannot write joint position for non-articulated entity.")
    jid = joint_ids if joint_ids is not None else slice(None)
    self.data.qpos[self._resolve_env_ids(env_ids), self.indexing.joint_q_adr[jid]] = position

  def write_joint_velocity(self, velocity, joint_ids=None, env_ids=None) -> None:
    if not self.is_articulated:
      raise ValueError("Cannot write joint velocity for non-articulated entity.")
    jid = joint_ids if joint_ids is not None else slice(None)
    self.data.qvel[self._resolve_env_ids(env_ids), self.indexing.joint_v_adr[jid]] = velocity

  def write_external_wrench(self, force, torque, body_ids: Sequence[int] | slice | None = None, env_ids=None) -> None:
    bid = body_ids if body_ids is not None else slice(None)
    g = self.indexing.body_ids[bid]
    e = self._resolve_env_ids(env_ids)
    if force is not None:
      self.data.xfrc_applied[e, g, 0:3] = force
    if torque is not None:
      self.data.xfrc_applied[e, g, 3:6] = torque

  def write_ctrl(self, ctrl, ctrl_ids=None, env_ids=None) -> None:
    if not self.is_actuated:
      raise ValueError("Cannot write control for non-actuated entity.")
    cid = ctrl_ids if ctrl_ids is not None else slice(None)
    self.data.ctrl[self._resolve_env_ids(env_ids), self.indexing.ctrl_ids[cid]] = ctrl

  def clear_state(self, env_ids=None) -> None:
    e = self._resolve_env_ids(env_ids)
    if not self.is_fixed_base:
      self.data.qfrc_applied[e, self.indexing.free_joint_v_adr] = 0.0
    self.data.xfrc_applied[e, self.indexing.body_ids] = 0.0
    if self.is_actuated:
      self.data.ctrl[e, self.indexing.ctrl_ids] = 0.0

  # -- root (data.py:190-240) -----------------------------------------------------------------------
  @property
  def _rb(self):
    return self.indexing.root_body_id

  @property
  def root_link_pose_w(self):
    return torch.cat([self.data.xpos[:, self._rb], self.data.xquat[:, self._rb]], dim=-1)

  # Velocities and body-frame quantities: when the data struct carries the engine's per-body outputs
  # (``link_vel_w`` / ``com_vel_w`` / ``link_state_b``, written by the kinematics phase of the step kernel,
  # SURVEY.md §8f-2) the properties are plain gathers; otherwise (``native=False``, or a data struct from another
  # engine) they are computed from ``cvel`` as the reference does (``data.py:20-31,472-516``).
  @property
  def _native(self) -> bool:
    return self.native and hasattr(self.data, "link_state_b")

  @property
  def root_link_vel_w(self):
    if self._native:
      return self.data.link_vel_w[:, self._rb]
    return compute_velocity_from_cvel(self.data.xpos[:, self._rb], self.data.subtree_com[:, self._rb],
                                      self.data.cvel[:, self._rb])

  @property
  def root_com_pose_w(self):
    q = quat_mul(self.data.xquat[:, self._rb], self.model.body_iquat[:, self._rb])
    return torch.cat([self.data.xipos[:, self._rb], q], dim=-1)

  @property
  def root_com_vel_w(self):
    if self._native:
      return self.data.com_vel_w[:, self._rb]
    return compute_velocity_from_cvel(self.data.xipos[:, self._rb], self.data.subtree_com[:, self._rb],
                                      self.data.cvel[:, self._rb])

  # -- bodies / geoms / sites (data.py:242-330) -------------------------------------------------------
  @property
  def body_link_pose_w(self):
    b = self.indexing.body_ids
    return torch.cat([self.data.xpos[:, b], self.data.xquat[:, b]], dim=-1)

  @property
  def body_link_vel_w(self):
    b = self.indexing.body_ids
    if self._native:
      return self.data.link_vel_w[:, b]
    return compute_velocity_from_cvel(self.data.xpos[:, b], self.data.subtree_com[:, self._rb].unsqueeze(1),
                                      self.data.cvel[:, b])

  @property
  def body_com_pose_w(self):
    b = self.indexing.body_ids
    return torch.cat([self.data.xipos[:, b], quat_mul(self.data.xquat[:, b], self.model.body_iquat[:, b])], dim=-1)

  @property
  def body_com_vel_w(self):
    b = self.indexing.body_ids
    if self._native:
      return self.data.com_vel_w[:, b]
    return compute_velocity_from_cvel(self.data.xipos[:, b], self.data.subtree_com[:, self._rb].unsqueeze(1),
                                      self.data.cvel[:, b])

  @property
  def body_external_wrench(self):
    return self.data.xfrc_applied[:, self.indexing.body_ids]

  @property
  def geom_pose_w(self):
    g = self.indexing.geom_ids
    return torch.cat([self.data.geom_xpos[:, g], quat_from_matrix(self.data.geom_xmat[:, g])], dim=-1)

  @property
  def geom_vel_w(self):
    g = self.indexing.geom_ids
    bodies = self.model.geom_bodyid[g]
    return compute_velocity_from_cvel(self.data.geom_xpos[:, g], self.data.subtree_com[:, self._rb].unsqueeze(1),
                                      self.data.cvel[:, bodies])

  @property
  def site_pose_w(self):
    s = self.indexing.site_ids
    return torch.cat([self.data.site_xpos[:, s], quat_from_matrix(self.data.site_xmat[:, s])], dim=-1)

  @property
  def site_vel_w(self):
    s = self.indexing.site_ids
    bodies = self.model.site_bodyid[s]
    return compute_velocity_from_cvel(self.data.site_xpos[:, s], self.data.subtree_com[:, self._rb].unsqueeze(1),
                                      self.data.cvel[:, bodies])

  # -- joints / actuators / sensors (data.py:332-352) --------------------------------------------------
  @property
  def joint_pos(self):
    return self.data.qpos[:, self.indexing.joint_q_adr]

  @property
  def joint_vel(self):
    return self.data.qvel[:, self.indexing.joint_v_adr]

  @property
  def joint_acc(self):
    return self.data.qacc[:, self.indexing.joint_v_adr]

  @property
  def actuator_force(self):
    return self.data.actuator_force[:, self.indexing.ctrl_ids]

  @property
  def generalized_force(self):
    return self.data.qfrc_applied[:, self.indexing.free_joint_v_adr]

  @property
  def sensor_data(self) -> dict:
    return {n: self.data.sensordata[:, idx] for n, idx in self.indexing.sensor_adr.items()}

  # -- sliced views (data.py:354-470) --------------------------------------------------------------------
  root_link_pos_w = property(lambda s: s.root_link_pose_w[:, 0:3])
  root_link_quat_w = property(lambda s: s.root_link_pose_w[:, 3:7])
  root_link_lin_vel_w = property(lambda s: s.root_link_vel_w[:, 0:3])
  root_link_ang_vel_w = property(lambda s: s.root_link_vel_w[:, 3:6])
  root_com_pos_w = property(lambda s: s.root_com_pose_w[:, 0:3])
  root_com_quat_w = property(lambda s: s.root_com_pose_w[:, 3:7])
  root_com_lin_vel_w = property(lambda s: s.root_com_vel_w[:, 0:3])
  root_com_ang_vel_w = property(lambda s: s.root_com_vel_w[:, 3:6])
  body_link_pos_w = property(lambda s: s.body_link_pose_w[..., 0:3])
  body_link_quat_w = property(lambda s: s.body_link_pose_w[..., 3:7])
  body_link_lin_vel_w = property(lambda s: s.body_link_vel_w[..., 0:3])
  body_link_ang_vel_w = property(lambda s: s.body_link_vel_w[..., 3:6])
  body_com_pos_w = property(lambda s: s.body_com_pose_w[..., 0:3])
  body_com_quat_w = property(lambda s: s.body_com_pose_w[..., 3:7])
  body_com_lin_vel_w = property(lambda s: s.body_com_vel_w[..., 0:3])
  body_com_ang_vel_w = property(lambda s: s.body_com_vel_w[..., 3:6])
  body_external_force = property(lambda s: s.body_external_wrench[..., 0:3])
  body_external_torque = property(lambda s: s.body_external_wrench[..., 3:6])
  geom_pos_w = property(lambda s: s.geom_pose_w[..., 0:3])
  geom_quat_w = property(lambda s: s.geom_pose_w[..., 3:7])
  geom_lin_vel_w = property(lambda s: s.geom_vel_w[..., 0:3])
  geom_ang_vel_w = property(lambda s: s.geom_vel_w[..., 3:6])
  site_pos_w = property(lambda s: s.site_pose_w[..., 0:3])
  site_quat_w = property(lambda s: s.site_pose_w[..., 3:7])
  site_lin_vel_w = property(lambda s: s.site_vel_w[..., 0:3])
  site_ang_vel_w = property(lambda s: s.site_vel_w[..., 3:6])

  # -- derived (data.py:472-516) ---------------------------------------------------------------------------
  @property
  def projected_gravity_b(self):
    if self._native:
      return self.data.link_state_b[:, self._rb, 6:9]
    return quat_apply_inverse(self.root_link_quat_w, self.gravity_vec_w)

  @property
  def heading_w(self):
    if self._native:
      return self.data.link_state_b[:, self._rb, 9]
    fwd = quat_apply(self.root_link_quat_w, self.forward_vec_b)
    return torch.atan2(fwd[:, 1], fwd[:, 0])

  @property
  def root_link_lin_vel_b(self):
    if self._native:
      return self.data.link_state_b[:, self._rb, 0:3]
    return quat_apply_inverse(self.root_link_quat_w, self.root_link_lin_vel_w)

  @property
  def root_link_ang_vel_b(self):
    if self._native:
      return self.data.link_state_b[:, self._rb, 3:6]
    return quat_apply_inverse(self.root_link_quat_w, self.root_link_ang_vel_w)

  @property
  def root_com_lin_vel_b(self):
    return quat_apply_inverse(self.root_link_quat_w, self.root_com_lin_vel_w)

  @property
  def root_com_ang_vel_b(self):
    return quat_apply_inverse(self.root_link_quat_w, self.root_com_ang_vel_w)
