"""EntityData / EntityIndexing on CPU tensors produced by the oracle.

Pins the frame conventions the reference relies on: the numeric self-check of
``scripts/csv_to_npz.py:279-284`` (root link velocities equal the commanded base velocity), exact landing of
written root/joint state in ``qpos/qvel`` (``tests/test_entity.py:269-290``), finite-difference body
velocities."""

from types import SimpleNamespace

import numpy as np
import pytest
import torch

from mjlab_b200.entity_data import EntityData, EntityIndexing, quat_from_matrix, quat_mul
from oracle.oracle import Oracle
from util import load_oracle, make_states


def _fake_sim(model, o, n):
  nb, ng, ns = int(model.nbody), int(model.ngeom), int(model.nsite)

  def t(x, *shape):
    return torch.tensor(np.array(x), dtype=torch.float32).reshape(n, *shape)

  data = SimpleNamespace(
    qpos=t(o.qpos, -1), qvel=t(o.qvel, -1), qacc=t(o.qacc, -1), ctrl=t(o.ctrl, -1),
    qfrc_applied=t(o.qfrc_applied, -1), xfrc_applied=t(o.xfrc_applied, nb, 6),
    xpos=t(o.xpos, nb, 3), xquat=t(o.xquat, nb, 4), xipos=t(o.xipos, nb, 3),
    subtree_com=t(o.subtree_com, nb, 3), cvel=t(o.cvel, nb, 6), geom_xpos=t(o.geom_xpos, ng, 3),
    geom_xmat=t(o.geom_xmat, ng, 3, 3), site_xpos=t(o.site_xpos, ns, 3), site_xmat=t(o.site_xmat, ns, 3, 3),
    sensordata=t(o.sensordata, -1), actuator_force=t(o.actuator_force, -1),
  )
  mdl = SimpleNamespace(
    body_iquat=torch.tensor(model.body_iquat, dtype=torch.float32)[None],
    jnt_range=torch.tensor(model.jnt_range, dtype=torch.float32)[None],
    geom_bodyid=torch.tensor(model.geom_bodyid, dtype=torch.int), site_bodyid=torch.tensor(model.site_bodyid, dtype=torch.int),
  )
  return data, mdl


@pytest.fixture(scope="module")
def g1_entity(g1_model):
  n = 6
  o = Oracle(g1_model, nworld=n)
  st = make_states(g1_model, n, seed=31, vel=1.0)
  load_oracle(o, st)
  o.forward()
  data, mdl = _fake_sim(g1_model, o, n)
  ix = EntityIndexing.from_model(g1_model, "robot")
  return EntityData(ix, data, mdl, "cpu", n, soft_joint_pos_limit_factor=0.9), o, st, n


def test_indexing_matches_compiled_ids(g1_model):
  ix = EntityIndexing.from_model(g1_model, "robot")
  assert ix.root_body_id == 2 and ix.body_names[0] == "pelvis" and len(ix.body_names) == 30
  assert ix.free_joint_q_adr.tolist() == list(range(7)) and ix.free_joint_v_adr.tolist() == list(range(6))
  assert ix.joint_q_adr.tolist() == list(range(7, 36)) and ix.joint_v_adr.tolist() == list(range(6, 35))
  assert ix.ctrl_ids.tolist() == list(range(29)) and ix.joint_names[0] == "left_hip_pitch_joint"
  assert ix.body_ids.dtype == torch.int32 and len(ix.geom_names) == 68 and "terrain" not in ix.geom_names
  assert set(ix.sensor_adr) == {"left_foot_ground_contact", "right_foot_ground_contact"}
  with pytest.raises(ValueError):
    EntityIndexing.from_model(g1_model, "nobody")


def test_root_velocity_conventions(g1_entity):
  ed, o, st, n = g1_entity
  qvel = torch.tensor(st["qvel"], dtype=torch.float32)
  # free joint: qvel[0:3] is the world-frame linear velocity of the root link origin,
  # qvel[3:6] the body-frame angular velocity (csv_to_npz.py:279-284)
  assert torch.allclose(ed.root_link_lin_vel_w, qvel[:, 0:3], atol=1e-5)
  assert torch.allclose(ed.root_link_ang_vel_b, qvel[:, 3:6], atol=1e-5)
  assert torch.allclose(ed.root_link_pos_w, torch.tensor(st["qpos"][:, 0:3], dtype=torch.float32), atol=1e-6)
  assert torch.allclose(ed.joint_pos, torch.tensor(st["qpos"][:, 7:], dtype=torch.float32))
  assert ed.projected_gravity_b.shape == (n, 3) and torch.allclose(ed.projected_gravity_b.norm(dim=1), torch.ones(n))
  assert ed.root_link_pose_w.shape == (n, 7) and ed.body_link_vel_w.shape == (n, 30, 6)


def test_body_velocities_match_finite_differences(g1_entity, g1_model):
  ed, o, st, n = g1_entity
  h = 1e-6
  o2 = Oracle(g1_model, nworld=n)
  qpos = st["qpos"].copy()
  qvel = st["qvel"]
  qpos[:, 0:3] += h * qvel[:, 0:3]
  w = qvel[:, 3:6]  # body-frame angular velocity: q <- q * exp(h w / 2)
  dq = np.concatenate([np.ones((n, 1)), 0.5 * h * w], axis=1)
  q = torch.tensor(qpos[:, 3:7])
  qpos[:, 3:7] = quat_mul(q, torch.tensor(dq)).numpy()
  qpos[:, 7:] += h * qvel[:, 6:]
  o2.qpos[:] = qpos
  o2.forward()
  nb = int(g1_model.nbody)
  fd_link = (o2.xpos.reshape(n, nb, 3) - o.xpos.reshape(n, nb, 3))[:, 2:] / h
  fd_com = (o2.xipos.reshape(n, nb, 3) - o.xipos.reshape(n, nb, 3))[:, 2:] / h
  assert np.abs(ed.body_link_lin_vel_w.numpy() - fd_link).max() < 2e-3
  assert np.abs(ed.body_com_lin_vel_w.numpy() - fd_com).max() < 2e-3
  ng = int(g1_model.ngeom)
  fd_geom = (o2.geom_xpos.reshape(n, ng, 3) - o.geom_xpos.reshape(n, ng, 3))[:, 1:] / h
  assert np.abs(ed.geom_lin_vel_w.numpy() - fd_geom).max() < 2e-3


def test_pose_accessors(g1_entity, g1_model):
  ed, o, st, n = g1_entity
  ng = int(g1_model.ngeom)
  # geom quaternion from its rotation matrix == xquat(body) * geom_quat
  gq = ed.geom_quat_w
  xq = torch.tensor(np.array(o.xquat), dtype=torch.float32).reshape(n, -1, 4)
  body = torch.tensor(g1_model.geom_bodyid[1:], dtype=torch.long)
  ref = quat_mul(xq[:, body], torch.tensor(g1_model.geom_quat[1:], dtype=torch.float32)[None])
  ref = torch.where(ref[..., 0:1] < 0, -ref, ref)
  assert gq.shape == (n, ng - 1, 4) and torch.allclose(gq, ref, atol=2e-5)
  m = torch.eye(3).expand(4, 3, 3)
  assert torch.allclose(quat_from_matrix(m), torch.tensor([1.0, 0, 0, 0]).expand(4, 4))
  assert ed.site_pose_w.shape == (n, 6, 7) and ed.body_com_pose_w.shape == (n, 30, 7)
  assert ed.heading_w.shape == (n,)
  assert set(ed.sensor_data) == {"left_foot_ground_contact", "right_foot_ground_contact"}
  lo, hi = ed.soft_joint_pos_limits[0, :, 0], ed.soft_joint_pos_limits[0, :, 1]
  rng = torch.tensor(g1_model.jnt_range[1:], dtype=torch.float32)
  assert torch.allclose(hi - lo, 0.9 * (rng[:, 1] - rng[:, 0]), atol=1e-6)


def test_writers_land_exactly(g1_entity):
  ed, o, st, n = g1_entity
  env_ids = torch.tensor([1, 4])
  state = torch.arange(26, dtype=torch.float32).reshape(2, 13)
  ed.write_root_state(state, env_ids)
  assert torch.equal(ed.data.qpos[env_ids][:, :7], state[:, :7]) and torch.equal(ed.data.qvel[env_ids][:, :6], state[:, 7:])
  assert not torch.equal(ed.data.qpos[0, :7], state[0, :7])
  jp = torch.full((2, 3), 0.25)
  ed.write_joint_position(jp, joint_ids=torch.tensor([0, 5, 28]), env_ids=env_ids)
  assert torch.equal(ed.data.qpos[env_ids][:, [7, 12, 35]], jp)
  ed.write_ctrl(torch.ones(n, 29))
  assert torch.equal(ed.data.ctrl, torch.ones(n, 29))
  ed.write_external_wrench(torch.ones(2, 1, 3), 2 * torch.ones(2, 1, 3), body_ids=[3], env_ids=env_ids)
  g = int(ed.indexing.body_ids[3])
  assert torch.equal(ed.data.xfrc_applied[env_ids][:, g], torch.tensor([[1.0, 1, 1, 2, 2, 2]] * 2))
  assert torch.equal(ed.body_external_force[env_ids][:, 3], torch.ones(2, 3))
  ed.clear_state(env_ids)
  assert (ed.data.xfrc_applied[env_ids] == 0).all() and (ed.data.ctrl[env_ids] == 0).all()
  assert (ed.data.ctrl[0] == 1).all()


def test_writers_with_partial_slice(g1_entity):
  """env_ids given as a slice selects that range only (reference entity/data.py:180-188 passes slices through)."""
  ed, o, st, n = g1_entity
  before = ed.data.qpos.clone()
  pose = torch.tensor([[1.0, 2.0, 3.0, 1.0, 0.0, 0.0, 0.0]]).repeat(2, 1)
  ed.write_root_pose(pose, env_ids=slice(2, 4))
  assert torch.equal(ed.data.qpos[2:4, :7], pose)
  assert torch.equal(ed.data.qpos[:2], before[:2]) and torch.equal(ed.data.qpos[4:], before[4:])
  ed.write_ctrl(torch.ones(n, 29))
  ed.clear_state(env_ids=slice(0, 1))
  assert (ed.data.ctrl[0] == 0).all() and (ed.data.ctrl[1:] == 1).all()
