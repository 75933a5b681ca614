"""Slim motion-tracking flat-terrain environment (BASELINE.json configs[2]): the
*caller* of the hot path for the task with self-collision sensing and the higher contact count.

Follows ``ManagerBasedRlEnv.step`` (``src/mjlab/envs/manager_based_rl_env.py:106-147``) with the terms of
``tasks/tracking/tracking_env_cfg.py`` / ``tasks/tracking/config/g1/flat_env_cfg.py``: action -> 4 x {ctrl,
``sim.step``} -> terminations (time-out, anchor height / orientation, end-effector height) -> rewards (six
tracking terms, action rate, joint limits, self-collision count) -> reset of finished envs onto the motion
clip (RSI: clip frame + pose / velocity / joint noise, ``mdp/commands.py:283-352``) + forward -> command update
(``commands.py:354-388``: time step, anchor-relative body targets) -> interval push -> observations (policy
group with noise, critic group privileged).

The reference needs a motion ``.npz`` from a W&B artifact, which is unavailable offline; the clip here is the
synthetic one SURVEY.md §8d specifies: the KNEES_BENT pose held for 500 frames (same keys and shapes as
``commands.py:34-50``, indexed per step exactly like a real clip).  Sampling of the start frame is uniform (the
adaptive failure-bin sampler of ``commands.py:239-281`` is host-side bookkeeping and, on a static clip, has no
effect on the physics load).  Everything is mask-based torch code, so one env step is CUDA-graph capturable.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from mjlab_b200.asset_zoo import g1, load_compiled
from mjlab_b200.entity_data import EntityData, EntityIndexing, quat_apply, quat_apply_inverse, quat_mul
from mjlab_b200.envs.velocity_env import _resolve
from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg, native

VELOCITY_RANGE = ((-0.5, 0.5), (-0.5, 0.5), (-0.2, 0.2), (-0.52, 0.52), (-0.52, 0.52), (-0.78, 0.78))  # tracking_env_cfg.py:30-37
POSE_RANGE = ((-0.05, 0.05), (-0.05, 0.05), (-0.01, 0.01), (-0.1, 0.1), (-0.1, 0.1), (-0.2, 0.2))       # :61-68
BODY_NAMES = ("pelvis", "left_hip_roll_link", "left_knee_link", "left_ankle_roll_link", "right_hip_roll_link",
              "right_knee_link", "right_ankle_roll_link", "torso_link", "left_shoulder_roll_link", "left_elbow_link",
              "left_wrist_yaw_link", "right_shoulder_roll_link", "right_elbow_link", "right_wrist_yaw_link")  # g1/flat_env_cfg.py:26-41
EE_NAMES = ("left_ankle_roll_link", "right_ankle_roll_link", "left_wrist_yaw_link", "right_wrist_yaw_link")  # :48-53


@dataclass
class TrackingEnvCfg:
  num_envs: int = 4096
  decimation: int = 4                 # tracking_env_cfg.py:305
  episode_length_s: float = 10.0      # :306
  clip_frames: int = 500              # synthetic static clip
  push_interval_s: tuple = (1.0, 3.0)
  friction_range: tuple = (0.3, 1.2)
  joint_position_range: tuple = (-0.1, 0.1)
  env_spacing: float = 2.5
  seed: int = 42
  sim: SimulationCfg = field(
    default_factory=lambda: SimulationCfg(
      nconmax=150_000, njmax=250, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20)))  # :283-291


def quat_inv(q):
  return torch.cat([q[..., 0:1], -q[..., 1:4]], dim=-1)


def yaw_quat(q):
  w, x, y, z = q.unbind(-1)
  yaw = torch.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
  out = torch.zeros_like(q)
  out[..., 0] = torch.cos(yaw / 2)
  out[..., 3] = torch.sin(yaw / 2)
  return out


def quat_from_euler_xyz(r, p, y):
  cy, sy, cr, sr, cp, sp = torch.cos(y / 2), torch.sin(y / 2), torch.cos(r / 2), torch.sin(r / 2), torch.cos(p / 2), torch.sin(p / 2)
  return torch.stack([cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                      sy * cr * cp - cy * sr * sp], dim=-1)


def quat_error_magnitude(a, b):
  d = quat_mul(a, quat_inv(b))
  return 2.0 * torch.atan2(d[..., 1:4].norm(dim=-1), d[..., 0].abs())


def rot6(q):
  """First two columns of R(q), flattened row-major over (3, 2) (``matrix_from_quat(q)[..., :2]``)."""
  w, x, y, z = q.unbind(-1)
  c0 = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)], dim=-1)
  c1 = torch.stack([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)], dim=-1)
  return torch.stack([c0, c1], dim=-1).flatten(-2)


class TrackingFlatEnv:
  def __init__(self, cfg: TrackingEnvCfg, device: str = "cuda:0", model=None, native_mdp: bool = True):
    self.cfg, self.device = cfg, device
    self.native_mdp = native_mdp  # fused CUDA kernels (csrc/b2_trackenv.cuh) or the torch reference below
    self.model = m = model if model is not None else load_compiled("g1_tracking_flat")
    n = self.num_envs = cfg.num_envs
    self.sim = Simulation(n, cfg.sim, m, device)
    dev = torch.device(device)
    f32 = dict(dtype=torch.float32, device=dev)
    self.nq, self.nv, self.nu = int(m.nq), int(m.nv), int(m.nu)
    self.step_dt = cfg.decimation * float(m.opt_timestep)
    self.max_episode_length = math.ceil(cfg.episode_length_s / self.step_dt)
    self.gen = torch.Generator(device=dev)
    self.gen.manual_seed(cfg.seed)
    key = m.keys["robot/init_state"]
    self.default_qpos = torch.tensor(key["qpos"], **f32)
    self.default_joint_pos = self.default_qpos[7:].clone()
    ix = EntityIndexing.from_model(m, "robot", device)
    self.robot = EntityData(ix, self.sim.data, self.sim.model, device, n, soft_joint_pos_limit_factor=0.9)
    self.action_scale = torch.tensor(_resolve(g1.ACTION_SCALE, list(ix.joint_names), 0.5), **f32)
    self.body_idx = torch.tensor([ix.body_names.index(b) for b in BODY_NAMES], device=dev)
    self.anchor = BODY_NAMES.index("torso_link")
    self.ee_idx = torch.tensor([BODY_NAMES.index(b) for b in EE_NAMES], device=dev)
    self.self_collision_adr = [int(v) for v in ix.sensor_adr["self_collision"]]  # python ints: graph-capturable indexing
    cols = math.ceil(math.sqrt(n))
    idx = torch.arange(n, device=dev)
    self.env_origins = torch.stack(
      [(idx // cols - (cols - 1) / 2) * cfg.env_spacing, (idx % cols - (cols - 1) / 2) * cfg.env_spacing,
       torch.zeros(n, device=dev)], dim=1).float()
    # ---- startup domain randomisation (tracking_env_cfg.py:162-198) -----------------------------------
    self.sim.expand_model_fields(["body_ipos", "qpos0", "geom_friction"])
    torso = ix.body_ids[ix.body_names.index("torso_link")].long()
    rng = torch.tensor([[-0.025, 0.025], [-0.05, 0.05], [-0.05, 0.05]], **f32)
    self.sim.model.body_ipos[:, torso] += self._rand(n, 3) * (rng[:, 1] - rng[:, 0]) + rng[:, 0]
    self.sim.model.qpos0[:, ix.joint_q_adr.long()] += self._rand(n, self.nu) * 0.02 - 0.01
    feet = torch.tensor([m.names["geom"].index(f"robot/{g}") for g in g1.FOOT_GEOMS], device=dev)
    lo, hi = cfg.friction_range
    self.sim.model.geom_friction[:, feet, 0] = self._rand(n, len(feet)) * (hi - lo) + lo
    # ---- synthetic clip: the keyframe held static (keys of commands.py:34-50) --------------------------
    self.sim.data.qpos[:] = self.default_qpos
    self.sim.data.qvel[:] = 0.0
    self.sim.forward()
    T, nbk = cfg.clip_frames, len(BODY_NAMES)
    pose0 = self.robot.body_link_pose_w[0:1, self.body_idx].clone()  # (1, 14, 7) at the world origin
    self.motion = dict(
      joint_pos=self.default_joint_pos.expand(T, -1).contiguous(), joint_vel=torch.zeros(T, self.nu, **f32),
      body_pos_w=pose0[:, :, 0:3].expand(T, -1, -1).contiguous(), body_quat_w=pose0[:, :, 3:7].expand(T, -1, -1).contiguous(),
      body_lin_vel_w=torch.zeros(T, nbk, 3, **f32), body_ang_vel_w=torch.zeros(T, nbk, 3, **f32))
    self.time_steps = torch.zeros(n, dtype=torch.long, device=dev)
    self.episode_length_buf = torch.zeros(n, dtype=torch.int32, device=dev)
    self.last_action = torch.zeros(n, self.nu, **f32)
    self.push_time_left = torch.zeros(n, **f32)
    lo_t, hi_t = cfg.push_interval_s
    self.push_time_left.copy_(self._rand(n) * (hi_t - lo_t) + lo_t)
    self.body_pos_relative_w = torch.zeros(n, nbk, 3, **f32)
    self.body_quat_relative_w = torch.zeros(n, nbk, 4, **f32)
    self._done_buf = torch.zeros(n, dtype=torch.bool, device=dev)
    self.log_row = torch.zeros(n, 3, **f32)  # (reward, terminated, truncated) for the rank-0 log gather
    self._vr = torch.tensor(VELOCITY_RANGE, **f32)
    self._pr = torch.tensor(POSE_RANGE, **f32)
    self._graph = None
    self.reset()

  # -- helpers -------------------------------------------------------------------------------------
  def _rand(self, *shape):
    return torch.rand(shape, generator=self.gen, device=self.device)

  def _clip(self, name):
    return self.motion[name][self.time_steps]

  def _resample(self, mask: torch.Tensor, restart_clip: torch.Tensor, Ub: torch.Tensor) -> None:
    """``MotionCommand._resample_command`` for the masked envs: new clip frame, RSI noise, state write.  ``Ub``:
    this call's uniforms ``[n, 13 + nu]`` (0 clip frame, 1-6 pose, 7-12 velocity, 13.. joints; include/b2sim.h)."""
    n, d = self.num_envs, self.sim.data
    T = self.cfg.clip_frames
    ts = (Ub[:, 0] * (T - 1)).long()
    self.time_steps.copy_(torch.where(restart_clip, ts, self.time_steps))
    pose_n = Ub[:, 1:7] * (self._pr[:, 1] - self._pr[:, 0]) + self._pr[:, 0]
    vel_n = Ub[:, 7:13] * (self._vr[:, 1] - self._vr[:, 0]) + self._vr[:, 0]
    root_pos = self._clip("body_pos_w")[:, 0] + self.env_origins + pose_n[:, 0:3]
    root_ori = quat_mul(quat_from_euler_xyz(pose_n[:, 3], pose_n[:, 4], pose_n[:, 5]), self._clip("body_quat_w")[:, 0])
    lin = self._clip("body_lin_vel_w")[:, 0] + vel_n[:, 0:3]
    ang = self._clip("body_ang_vel_w")[:, 0] + vel_n[:, 3:6]
    lo, hi = self.cfg.joint_position_range
    jp = self._clip("joint_pos") + Ub[:, 13:13 + self.nu] * (hi - lo) + lo
    lim = self.robot.soft_joint_pos_limits
    jp = torch.minimum(torch.maximum(jp, lim[..., 0]), lim[..., 1])
    qpos = torch.cat([root_pos, root_ori, jp], dim=1)
    qvel = torch.cat([lin, quat_apply_inverse(root_ori, ang), self._clip("joint_vel")], dim=1)
    mk = mask.unsqueeze(1)
    d.qpos[:] = torch.where(mk, qpos, d.qpos[:])
    d.qvel[:] = torch.where(mk, qvel, d.qvel[:])
    d.ctrl[:] = torch.where(mk, torch.zeros_like(d.ctrl[:]), d.ctrl[:])  # clear_state (data.py:170-178)
    d.xfrc_applied[:] = torch.where(mask.view(n, 1, 1), torch.zeros_like(d.xfrc_applied[:]), d.xfrc_applied[:])

  def _robot_bodies(self):
    pose = self.robot.body_link_pose_w[:, self.body_idx]
    vel = self.robot.body_link_vel_w[:, self.body_idx]
    return pose[..., 0:3], pose[..., 3:7], vel[..., 0:3], vel[..., 3:6]

  def _update_command(self) -> None:
    """``MotionCommand._update_command``: anchor-relative targets (yaw-aligned, height from the clip)."""
    pos, quat, _, _ = self._robot_bodies()
    nbk = len(BODY_NAMES)
    a_pos = (self._clip("body_pos_w")[:, self.anchor] + self.env_origins)[:, None, :].expand(-1, nbk, -1)
    a_quat = self._clip("body_quat_w")[:, self.anchor][:, None, :].expand(-1, nbk, -1)
    r_pos = pos[:, self.anchor][:, None, :].expand(-1, nbk, -1)
    r_quat = quat[:, self.anchor][:, None, :].expand(-1, nbk, -1)
    delta_pos = torch.cat([r_pos[..., 0:2], a_pos[..., 2:3]], dim=-1)
    delta_ori = yaw_quat(quat_mul(r_quat, quat_inv(a_quat)))
    # (in place: the fused kernels hold pointers to these two tensors)
    self.body_quat_relative_w.copy_(quat_mul(delta_ori, self._clip("body_quat_w")))
    self.body_pos_relative_w.copy_(delta_pos + quat_apply(delta_ori, self._clip("body_pos_w") + self.env_origins[:, None, :] - a_pos))

  def observations(self, Z: torch.Tensor | None = None):
    """(policy, critic) groups of tracking_env_cfg.py:88-150; uniform noise on the policy group from the draws
    ``Z[n, 15 + 2 nu]`` (anchor pos 3, anchor ori 6, base lin 3, base ang 3, joint pos, joint vel)."""
    d = self.sim.data
    if Z is None:
      Z = self._rand(self.num_envs, 15 + 2 * self.nu)
    pos, quat, _, _ = self._robot_bodies()
    a_pos = self._clip("body_pos_w")[:, self.anchor] + self.env_origins
    a_quat = self._clip("body_quat_w")[:, self.anchor]
    r_pos, r_quat = pos[:, self.anchor], quat[:, self.anchor]
    command = torch.cat([self._clip("joint_pos"), self._clip("joint_vel")], dim=1)
    anchor_pos_b = quat_apply_inverse(r_quat, a_pos - r_pos)
    anchor_ori_b = rot6(quat_mul(quat_inv(r_quat), a_quat))
    base_lin, base_ang = self.robot.root_link_lin_vel_b, self.robot.root_link_ang_vel_b
    jp, jv = d.qpos[:, 7:] - self.default_joint_pos, d.qvel[:, 6:]
    terms = [command, anchor_pos_b, anchor_ori_b, base_lin, base_ang, jp, jv, self.last_action]
    amp = (0.0, 0.25, 0.05, 0.5, 0.2, 0.01, 0.5, 0.0)
    noisy, k = [], 0
    for t, a in zip(terms, amp):
      if a == 0.0:
        noisy.append(t)
      else:
        noisy.append(t + (Z[:, k:k + t.shape[1]] * 2 - 1) * a)
        k += t.shape[1]
    rq = r_quat[:, None, :].expand(-1, len(BODY_NAMES), -1)
    body_pos_b = quat_apply_inverse(rq, pos - r_pos[:, None, :]).flatten(1)
    body_ori_b = rot6(quat_mul(quat_inv(rq), quat)).flatten(1)
    critic = torch.cat(terms[:3] + [body_pos_b, body_ori_b] + terms[3:], dim=1)
    return torch.cat(noisy, dim=1), critic

  # -- API -----------------------------------------------------------------------------------------
  def reset(self):
    all_ = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
    self._resample(all_, all_, self._rand(self.num_envs, 13 + self.nu))
    self.sim.forward()
    self._update_command()
    return self.observations()

  def enable_cuda_graph(self) -> None:
    dev = torch.device(self.device)
    self._action_buf = torch.zeros(self.num_envs, self.nu, device=dev)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for _ in range(2):
        self._step_impl(self._action_buf)
      g = torch.cuda.CUDAGraph()
      g.register_generator_state(self.gen)
      with torch.cuda.graph(g, stream=side):
        self._graph_out = self._step_impl(self._action_buf)
    torch.cuda.current_stream(dev).wait_stream(side)
    self._graph = g

  def step(self, action: torch.Tensor):
    if self._graph is not None:
      self._action_buf.copy_(action, non_blocking=True)
      self._graph.replay()
      return self._graph_out
    return self._step_impl(action)

  def _step_impl(self, action: torch.Tensor):
    n = self.num_envs
    U = self._rand(n, 48 + 4 * self.nu)  # B2_TRACKENV_NU: every random number of this env step (include/b2sim.h)
    if self.native_mdp:
      return self._step_native(action, U)
    return self._step_torch(action, U)

  def _native_setup(self) -> None:
    """Argument block of the fused kernels: device pointers of the env's own tensors (all persistent)."""
    n, nu, dev = self.num_envs, self.nu, torch.device(self.device)
    f32 = dict(dtype=torch.float32, device=dev)
    nb = len(BODY_NAMES)
    self._U = torch.zeros(n, 48 + 4 * nu, **f32)
    self._act = torch.zeros(n, nu, **f32)
    self._reward = torch.zeros(n, **f32)
    self._term = torch.zeros(n, dtype=torch.bool, device=dev)
    self._trunc = torch.zeros(n, dtype=torch.bool, device=dev)
    self._obs = torch.zeros(n, 5 * nu + 15, **f32)
    self._critic = torch.zeros(n, 5 * nu + 15 + 9 * nb, **f32)
    self._body_idx32 = self.robot.indexing.body_ids[self.body_idx].to(torch.int32).contiguous()
    self._ee_idx32 = self.ee_idx.to(torch.int32).contiguous()
    lim = self.robot.soft_joint_pos_limits
    self._soft_lo, self._soft_hi = lim[0, :, 0].contiguous(), lim[0, :, 1].contiguous()
    self.body_pos_relative_w = self.body_pos_relative_w.contiguous()
    self.body_quat_relative_w = self.body_quat_relative_w.contiguous()
    A = native.B2TrackEnvArgs()
    ptr = dict(action=self._act, U=self._U, default_joint_pos=self.default_joint_pos, soft_lo=self._soft_lo,
               soft_hi=self._soft_hi, env_origins=self.env_origins, m_joint_pos=self.motion["joint_pos"],
               m_joint_vel=self.motion["joint_vel"], m_body_pos=self.motion["body_pos_w"],
               m_body_quat=self.motion["body_quat_w"], m_body_lin=self.motion["body_lin_vel_w"],
               m_body_ang=self.motion["body_ang_vel_w"], body_idx=self._body_idx32, ee_idx=self._ee_idx32,
               time_steps=self.time_steps, episode_length=self.episode_length_buf, last_action=self.last_action,
               push_time_left=self.push_time_left, body_pos_rel=self.body_pos_relative_w,
               body_quat_rel=self.body_quat_relative_w, reward=self._reward, terminated=self._term,
               truncated=self._trunc, mask=self._done_buf, log_row=self.log_row, obs=self._obs, critic=self._critic)
    for k, t in ptr.items():
      assert t.is_contiguous(), k
      setattr(A, k, t.data_ptr())
    for k in range(6):
      A.pose_range[2 * k], A.pose_range[2 * k + 1] = POSE_RANGE[k]
      A.vel_range[2 * k], A.vel_range[2 * k + 1] = VELOCITY_RANGE[k]
    A.step_dt = self.step_dt
    A.jp_lo, A.jp_hi = self.cfg.joint_position_range
    A.push_lo, A.push_hi = self.cfg.push_interval_s
    A.nb, A.nee, A.anchor, A.T = nb, len(EE_NAMES), self.anchor, self.cfg.clip_frames
    A.self_collision_adr = self.self_collision_adr[0]
    A.root_body = int(self.robot.indexing.root_body_id)
    A.max_episode_length = self.max_episode_length
    self._native_args = A
    self._keep = ptr

  def _step_native(self, action: torch.Tensor, U: torch.Tensor):
    """Two fused kernels around the masked forward (csrc/b2_trackenv.cuh) instead of ~150 torch launches."""
    import ctypes

    if getattr(self, "_native_args", None) is None:
      self._native_setup()
    sim = self.sim
    st = sim._stream()
    self._U.copy_(U)
    self._act.copy_(action)
    native.check(sim._lib.b2_velenv_pre(sim._h, ctypes.c_void_p(self._act.data_ptr()),
                                        ctypes.c_void_p(self.default_joint_pos.data_ptr()),
                                        ctypes.c_void_p(self.action_scale.data_ptr()), st))
    sim.step_n(self.cfg.decimation)
    native.check(sim._lib.b2_trackenv_post1(sim._h, ctypes.byref(self._native_args), st))
    sim.forward(env_mask=self._done_buf)
    native.check(sim._lib.b2_trackenv_post2(sim._h, ctypes.byref(self._native_args), st))
    return self._obs, self._reward, self._term, self._trunc, {"critic": self._critic}

  def _step_torch(self, action: torch.Tensor, U: torch.Tensor):
    cfg, d, n, nu = self.cfg, self.sim.data, self.num_envs, self.nu
    S = 13 + nu
    d.ctrl[:] = self.default_joint_pos + self.action_scale * action  # JointPositionAction, use_default_offset
    self.sim.step_n(cfg.decimation)
    self.episode_length_buf += 1
    pos, quat, lin, ang = self._robot_bodies()
    a_pos = self._clip("body_pos_w")[:, self.anchor] + self.env_origins
    a_quat = self._clip("body_quat_w")[:, self.anchor]
    r_pos, r_quat = pos[:, self.anchor], quat[:, self.anchor]
    # terminations (tracking_env_cfg.py:253-278)
    truncated = self.episode_length_buf >= self.max_episode_length
    bad_z = (a_pos[:, 2] - r_pos[:, 2]).abs() > 0.25
    g = self.robot.gravity_vec_w
    bad_ori = (quat_apply_inverse(a_quat, g)[:, 2] - quat_apply_inverse(r_quat, g)[:, 2]).abs() > 0.8
    ee_err = (self.body_pos_relative_w[:, self.ee_idx, 2] - pos[:, self.ee_idx, 2]).abs()
    terminated = bad_z | bad_ori | (ee_err > 0.25).any(dim=1)
    # rewards (:200-250)
    r = 0.5 * torch.exp(-((a_pos - r_pos) ** 2).sum(-1) / 0.3**2)
    r = r + 0.5 * torch.exp(-quat_error_magnitude(a_quat, r_quat) ** 2 / 0.4**2)
    r = r + torch.exp(-((self.body_pos_relative_w - pos) ** 2).sum(-1).mean(-1) / 0.3**2)
    r = r + torch.exp(-(quat_error_magnitude(self.body_quat_relative_w, quat) ** 2).mean(-1) / 0.4**2)
    r = r + torch.exp(-((self._clip("body_lin_vel_w") - lin) ** 2).sum(-1).mean(-1) / 1.0**2)
    r = r + torch.exp(-((self._clip("body_ang_vel_w") - ang) ** 2).sum(-1).mean(-1) / 3.14**2)
    r = r - 0.1 * ((action - self.last_action) ** 2).sum(1)
    jp, lim = d.qpos[:, 7:], self.robot.soft_joint_pos_limits
    r = r - 10.0 * ((lim[..., 0] - jp).clamp(min=0) + (jp - lim[..., 1]).clamp(min=0)).sum(1)
    r = r - 10.0 * d.sensordata[:, self.self_collision_adr[0]]
    reward = r * self.step_dt
    self.last_action.copy_(action)
    done = terminated | truncated
    self.log_row.copy_(torch.stack([reward, terminated.float(), truncated.float()], dim=1))
    # reset onto the clip (RSI); then the clip's time step: envs that ran off its end restart on it.  Both rewrite
    # the whole state of their envs, so one forward over the union follows (the reference runs one per event)
    self._resample(done, done, U[:, 0:S])
    self.episode_length_buf.copy_(torch.where(done, torch.zeros_like(self.episode_length_buf), self.episode_length_buf))
    self.last_action.copy_(torch.where(done.unsqueeze(1), torch.zeros_like(self.last_action), self.last_action))
    self.time_steps += 1
    ended = self.time_steps >= cfg.clip_frames
    self._resample(ended, ended, U[:, S:2 * S])
    self._done_buf.copy_(done | ended)
    self.sim.forward(env_mask=self._done_buf)
    self._update_command()
    # interval push (push_by_setting_velocity with VELOCITY_RANGE, tracking_env_cfg.py:155-161)
    self.push_time_left -= self.step_dt
    push = self.push_time_left <= 0
    P = U[:, 2 * S:2 * S + 7]
    vel_w = self.robot.root_link_vel_w + P[:, 0:6] * (self._vr[:, 1] - self._vr[:, 0]) + self._vr[:, 0]
    vel = torch.cat([vel_w[:, 0:3], quat_apply_inverse(self.robot.root_link_quat_w, vel_w[:, 3:6])], dim=1)
    d.qvel[:, 0:6] = torch.where(push.unsqueeze(1), vel, d.qvel[:, 0:6])
    lo_t, hi_t = cfg.push_interval_s
    self.push_time_left.copy_(torch.where(push, P[:, 6] * (hi_t - lo_t) + lo_t, self.push_time_left))
    policy, critic = self.observations(U[:, 2 * S + 7:])
    return policy, reward, terminated, truncated, {"critic": critic}

  def close(self):
    self.sim.close()
