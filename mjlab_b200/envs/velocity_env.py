"""Slim velocity-tracking flat-terrain environment: the *caller* of the hot path.

Follows the control flow of the reference's ``ManagerBasedRlEnv.step``
(``src/mjlab/envs/manager_based_rl_env.py:106-147``) for the velocity task
(``tasks/velocity/velocity_env_cfg.py``): process action -> ``decimation`` x {write ctrl,
``sim.step``} -> terminations -> rewards -> reset terminated envs + ``sim.forward`` -> commands
-> interval push event -> observations.  The manager/term machinery is not rebuilt (out of scope,
SURVEY.md §2 rows 6-8); terms are fused torch expressions and resets are mask-based so the step
has no host synchronisation.
"""

from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import torch

from mjlab_b200.asset_zoo import g1, go1, load_compiled
from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg


@dataclass
class VelocityEnvCfg:
  robot: str = "g1"
  num_envs: int = 4096
  decimation: int = 4                      # velocity_env_cfg.py:271
  episode_length_s: float = 20.0           # :272
  fall_angle: float = math.radians(70.0)   # :241-243
  push_interval_s: tuple = (1.0, 3.0)      # :156-161
  push_vel: float = 0.5                    # flat_env_cfg.py:20-24
  friction_range: tuple = (0.3, 1.2)       # :162-172
  env_spacing: float = 2.5
  terrain: str = "flat"                    # "flat" (plane) | "rough" (box terrain, velocity_env_cfg.py:274-278)
  max_init_terrain_level: int = 5          # terrain_importer.py:203-223
  seed: int = 42
  sim: SimulationCfg = field(
    default_factory=lambda: SimulationCfg(
      nconmax=140_000, njmax=300,
      mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20),
    )
  )


def _resolve(pattern_map: dict, names: list[str], default: float) -> list[float]:
  out = []
  for n in names:
    for pat, v in pattern_map.items():
      if re.match(pat, n):
        out.append(v)
        break
    else:
      out.append(default)
  return out


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
  w, u = q[:, 0:1], q[:, 1:4]
  t = 2.0 * torch.cross(u, v, dim=-1)
  return v - w * t + torch.cross(u, t, dim=-1)


class VelocityFlatEnv:
  def __init__(self, cfg: VelocityEnvCfg, device: str = "cuda:0", model=None, native_mdp: bool = True):
    self.cfg = cfg
    self.device = device
    zoo = g1 if cfg.robot == "g1" else go1
    self.model = model if model is not None else load_compiled(f"{cfg.robot}_{cfg.terrain}")
    m = self.model
    self.num_envs = cfg.num_envs
    self.sim = Simulation(cfg.num_envs, cfg.sim, m, device)
    self.nq, self.nv, self.nu = int(m.nq), int(m.nv), int(m.nu)
    self.step_dt = cfg.decimation * float(m.opt_timestep)
    self.max_episode_length = math.ceil(cfg.episode_length_s / self.step_dt)
    dev = torch.device(device)
    self.gen = torch.Generator(device=dev)
    self.gen.manual_seed(cfg.seed)
    key = m.keys["robot/init_state"]
    f32 = dict(dtype=torch.float32, device=dev)
    self.default_qpos = torch.tensor(key["qpos"], **f32)
    self.default_joint_pos = self.default_qpos[7:].clone()
    jnames = [n.split("/")[-1] for n in m.names["joint"][1:]]
    self.action_scale = torch.tensor(_resolve(zoo.ACTION_SCALE, jnames, 0.5), **f32)
    rng = torch.tensor(m.jnt_range[1:], **f32)
    mid, half = rng.mean(dim=1), 0.5 * (rng[:, 1] - rng[:, 0]) * 0.9  # soft limits (entity.py:372-380)
    self.soft_lo, self.soft_hi = mid - half, mid + half
    # env origins on a grid (terrain_importer.py:203-223 for the flat case)
    n = cfg.num_envs
    cols = math.ceil(math.sqrt(n))
    idx = torch.arange(n, device=dev)
    self.env_origins = torch.stack(
      [(idx // cols - (cols - 1) / 2) * cfg.env_spacing, (idx % cols - (cols - 1) / 2) * cfg.env_spacing,
       torch.zeros(n, device=dev)], dim=1).float()
    if "terrain_origins" in m.arrays:
      # box terrain: spawn on the sub-terrain origins, level <= max_init_terrain_level, types round-robin
      # (terrain_importer.py:203-223); the curriculum's level updates are host-side MDP and not restated
      from mjlab_b200.terrains import env_origins_curriculum

      org, self.terrain_levels, self.terrain_types = env_origins_curriculum(
        n, m.arrays["terrain_origins"], cfg.max_init_terrain_level, cfg.seed)
      self.env_origins = torch.tensor(org, dtype=torch.float32, device=dev)
    # startup domain randomisation: foot friction (events.py:212-265 'abs' on geom_friction[:, feet, 0])
    self.sim.expand_model_fields(["geom_friction"])
    foot_ids = torch.tensor(
      [m.names["geom"].index(f"robot/{g}") for g in zoo.FOOT_GEOMS], device=dev, dtype=torch.long
    )
    lo, hi = cfg.friction_range
    mu = torch.rand((n, len(foot_ids)), generator=self.gen, device=dev) * (hi - lo) + lo
    self.sim.model.geom_friction[:, foot_ids, 0] = mu
    self.episode_length_buf = torch.zeros(n, dtype=torch.int32, device=dev)
    self.last_action = torch.zeros(n, self.nu, **f32)
    self.command = torch.zeros(n, 3, **f32)
    # UniformVelocityCommand state (velocity_env_cfg.py:66-83): resampling timer 3..8 s, heading control, 10 % standing
    self.cmd_time_left = torch.zeros(n, **f32)
    self.heading_target = torch.zeros(n, **f32)
    self.is_standing = torch.zeros(n, dtype=torch.bool, device=dev)
    self.log_row = torch.zeros(n, 3, **f32)  # (reward, terminated, truncated): the rank-0 log gather reads this
    self.push_time_left = torch.zeros(n, **f32)
    self._down = torch.tensor([0.0, 0.0, -1.0], **f32).expand(n, 3)  # gravity direction (world)
    self.native_mdp = bool(native_mdp)
    self._graph = None
    self._done_buf = torch.zeros(n, dtype=torch.bool, device=dev)  # fixed address (graph capture / native)
    lo_t, hi_t = cfg.push_interval_s
    self.push_time_left.copy_(self._rand(n) * (hi_t - lo_t) + lo_t)
    if self.native_mdp:
      self._init_native()
    self.reset()

  # -- helpers -----------------------------------------------------------------------------------
  def _rand(self, *shape):
    return torch.rand(shape, generator=self.gen, device=self.device)

  def _draw(self) -> torch.Tensor:
    """The step's uniform numbers, one launch (layout: include/b2sim.h B2VelEnvArgs.U): 0-1 reset xy, 2 yaw,
    3-5 command, 6 heading target, 7 standing draw, 8 command timer, 9-10 push, 11 push timer, 16.. obs noise."""
    return self._rand(self.num_envs, 16 + 9 + 2 * self.nu)

  def _init_native(self) -> None:
    from mjlab_b200.sim import native

    n, dev = self.num_envs, self.device
    self.nobs = 9 + 3 * self.nu + 3
    self._obs = torch.zeros(n, self.nobs, device=dev)
    self._reward = torch.zeros(n, device=dev)
    self._term = torch.zeros(n, dtype=torch.bool, device=dev)
    self._trunc = torch.zeros(n, dtype=torch.bool, device=dev)
    self._U = torch.zeros(n, 16 + 9 + 2 * self.nu, device=dev)
    self._critic = torch.zeros(n, self.nobs, device=dev)
    self._action_in = torch.zeros(n, self.nu, device=dev)
    self._origins = self.env_origins.contiguous()
    a = native.B2VelEnvArgs()
    for name, t in (
      ("action", self._action_in), ("U", self._U), ("default_qpos", self.default_qpos),
      ("default_joint_pos", self.default_joint_pos), ("action_scale", self.action_scale),
      ("soft_lo", self.soft_lo), ("soft_hi", self.soft_hi), ("env_origins", self._origins),
      ("episode_length", self.episode_length_buf), ("last_action", self.last_action),
      ("command", self.command), ("push_time_left", self.push_time_left), ("obs", self._obs),
      ("reward", self._reward), ("terminated", self._term), ("truncated", self._trunc), ("done", self._done_buf),
      ("cmd_time_left", self.cmd_time_left), ("heading_target", self.heading_target),
      ("is_standing", self.is_standing), ("critic", self._critic), ("log_row", self.log_row),
    ):
      assert t.is_contiguous()
      setattr(a, name, t.data_ptr())
    a.step_dt, a.fall_angle, a.push_vel = self.step_dt, self.cfg.fall_angle, self.cfg.push_vel
    a.push_lo, a.push_hi = self.cfg.push_interval_s
    a.max_episode_length = self.max_episode_length
    self._native_args = a

  def _reset_where(self, mask: torch.Tensor, U: torch.Tensor) -> None:
    """reset_root_state_uniform + reset_joints_by_scale (envs/mdp/events.py:43-124) for masked envs."""
    d = self.sim.data
    n = self.num_envs
    qpos = self.default_qpos.expand(n, -1).clone()
    qpos[:, 0:2] += (U[:, 0:2] - 0.5) + self.env_origins[:, 0:2]
    qpos[:, 2] += self.env_origins[:, 2]
    yaw = (U[:, 2] * 2 - 1) * 3.14
    qpos[:, 3] = torch.cos(0.5 * yaw)
    qpos[:, 6] = torch.sin(0.5 * yaw)
    qpos[:, 7:] = torch.minimum(torch.maximum(qpos[:, 7:], self.soft_lo), self.soft_hi)
    mk = mask.unsqueeze(1)
    # in-place updates everywhere: the step may be replayed from a CUDA graph (fixed addresses)
    d.qpos[:] = torch.where(mk, qpos, d.qpos[:])
    d.qvel[:] = torch.where(mk, torch.zeros_like(d.qvel[:]), d.qvel[:])
    d.ctrl[:] = torch.where(mk, self.default_joint_pos.expand(n, -1), d.ctrl[:])
    self.episode_length_buf.copy_(torch.where(mask, torch.zeros_like(self.episode_length_buf), self.episode_length_buf))
    self.last_action.copy_(torch.where(mk, torch.zeros_like(self.last_action), self.last_action))

  def _update_command(self, done: torch.Tensor, U: torch.Tensor) -> None:
    """CommandTerm.compute + UniformVelocityCommand (tasks/velocity/mdp/velocity_command.py:64-110): resample on
    reset and on timer expiry (3..8 s), heading control on every env, standing envs (10 %) get a zero command."""
    ctl = self.cmd_time_left - self.step_dt
    rs = done | (ctl <= 0)
    cmd = torch.stack([U[:, 3] * 2 - 1, U[:, 4] - 0.5, U[:, 5] * 2 - 1], dim=1)
    self.command.copy_(torch.where(rs.unsqueeze(1), cmd, self.command))
    self.heading_target.copy_(torch.where(rs, (U[:, 6] * 2 - 1) * math.pi, self.heading_target))
    self.is_standing.copy_(torch.where(rs, U[:, 7] <= 0.1, self.is_standing))
    self.cmd_time_left.copy_(torch.where(rs, 3.0 + 5.0 * U[:, 8], ctl))
    q = self.sim.data.qpos[:, 3:7]
    fx = 1 - 2 * (q[:, 2] ** 2 + q[:, 3] ** 2)
    fy = 2 * (q[:, 1] * q[:, 2] + q[:, 0] * q[:, 3])
    err = self.heading_target - torch.atan2(fy, fx)
    err = err - 2 * math.pi * torch.floor((err + math.pi) / (2 * math.pi))
    self.command[:, 2] = (0.5 * err).clamp(-1.0, 1.0)
    self.command.copy_(torch.where(self.is_standing.unsqueeze(1), torch.zeros_like(self.command), self.command))

  def observations(self, U: torch.Tensor | None = None):
    """Policy group (uniform noise per term, velocity_env_cfg.py:86-118); ``critic`` is the noise-free copy."""
    d = self.sim.data
    q = d.qpos[:, 3:7]
    lin_b = quat_rotate_inverse(q, d.qvel[:, 0:3])
    terms = [lin_b, d.qvel[:, 3:6], quat_rotate_inverse(q, self._down), d.qpos[:, 7:] - self.default_joint_pos,
             d.qvel[:, 6:], self.last_action, self.command]
    self.critic_obs = torch.cat(terms, dim=1)
    if U is None:
      return self.critic_obs
    Z, nu = U[:, 16:] * 2 - 1, self.nu
    noise = [Z[:, 0:3] * 0.1, Z[:, 3:6] * 0.2, Z[:, 6:9] * 0.05, Z[:, 9:9 + nu] * 0.01, Z[:, 9 + nu:9 + 2 * nu] * 1.5]
    return torch.cat([t + z for t, z in zip(terms[:5], noise)] + terms[5:], dim=1)

  # -- API ---------------------------------------------------------------------------------------
  def reset(self):
    all_ = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
    U = self._draw()
    self._reset_where(all_, U)
    self.sim.forward()
    self._update_command(all_, U)
    return self.observations()

  def enable_cuda_graph(self) -> None:
    """Capture one whole env step (ctrl write, 4 sub-steps, MDP glue, masked reset + forward) in a
    CUDA graph.  The step has no host synchronisation, so replay is a single launch; inputs and
    outputs live in static buffers (``step`` copies the action in and returns the static outputs)."""
    dev = torch.device(self.device)
    self._action_buf = torch.zeros(self.num_envs, self.nu, device=dev)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for _ in range(2):  # warm-up on the side stream (allocator, lazy inits)
        self._step_impl(self._action_buf)
      g = torch.cuda.CUDAGraph()
      g.register_generator_state(self.gen)  # torch.rand(generator=self.gen) inside the capture
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
    return self._step_native(action) if self.native_mdp else self._step_torch(action)

  def _step_native(self, action: torch.Tensor):
    """Two fused kernels (b2_velenv_pre / b2_velenv_post) instead of ~90 torch launches."""
    import ctypes

    from mjlab_b200.sim import native

    sim = self.sim
    self._action_in.copy_(action)
    self._U.copy_(self._draw())
    st = sim._stream()
    native.check(sim._lib.b2_velenv_pre(
      sim._h, ctypes.c_void_p(self._action_in.data_ptr()), ctypes.c_void_p(self.default_joint_pos.data_ptr()),
      ctypes.c_void_p(self.action_scale.data_ptr()), st))
    sim.step_n(self.cfg.decimation)
    native.check(sim._lib.b2_velenv_post(sim._h, ctypes.byref(self._native_args), st))
    sim.forward(env_mask=self._done_buf)
    return self._obs, self._reward, self._term, self._trunc, {"critic": self._critic}

  def _step_torch(self, action: torch.Tensor):
    """Reference implementation of the same step in torch ops (also the checker of the fused kernels)."""
    cfg, d = self.cfg, self.sim.data
    U = self._draw()
    # JointPositionAction: target = default + scale * action (joint_actions.py:85-103)
    d.ctrl[:] = self.default_joint_pos + self.action_scale * action
    self.sim.step_n(cfg.decimation)
    self.episode_length_buf += 1
    q = d.qpos[:, 3:7]
    grav_b = quat_rotate_inverse(q, self._down)
    terminated = torch.acos((-grav_b[:, 2]).clamp(-1.0, 1.0)) > cfg.fall_angle  # bad_orientation
    truncated = self.episode_length_buf >= self.max_episode_length              # time_out
    # rewards (velocity_env_cfg.py:183-214): tracking, posture, limits, action rate
    lin_b = quat_rotate_inverse(q, d.qvel[:, 0:3])
    r_lin = torch.exp(-((self.command[:, :2] - lin_b[:, :2]) ** 2).sum(1) / 0.25)
    r_ang = torch.exp(-((self.command[:, 2] - d.qvel[:, 5]) ** 2) / 0.25)
    jp = d.qpos[:, 7:]
    r_pose = torch.exp(-((jp - self.default_joint_pos) ** 2).mean(1) / 0.09)
    r_lim = -((self.soft_lo - jp).clamp(min=0) + (jp - self.soft_hi).clamp(min=0)).sum(1)
    r_rate = -0.1 * ((action - self.last_action) ** 2).sum(1)
    reward = (r_lin + r_ang + r_pose + r_lim + r_rate) * self.step_dt
    self.last_action.copy_(action)
    done = terminated | truncated
    # partial reset + forward (manager_based_rl_env.py:128-132), mask-based: no host sync
    self._reset_where(done, U)
    self._done_buf.copy_(done)
    self.sim.forward(env_mask=self._done_buf)  # only the reset envs need new derived quantities here
    self._update_command(done, U)
    self.log_row.copy_(torch.stack([reward, terminated.float(), truncated.float()], dim=1))
    # interval event: push_by_setting_velocity (events.py:127-143)
    self.push_time_left -= self.step_dt
    push = self.push_time_left <= 0
    pv = (U[:, 9:11] * 2 - 1) * cfg.push_vel
    d.qvel[:, 0:2] = torch.where(push.unsqueeze(1), pv, d.qvel[:, 0:2])
    lo_t, hi_t = cfg.push_interval_s
    self.push_time_left.copy_(torch.where(push, U[:, 11] * (hi_t - lo_t) + lo_t, self.push_time_left))
    obs = self.observations(U)
    return obs, reward, terminated, truncated, {"critic": self.critic_obs}

  def close(self):
    self.sim.close()
