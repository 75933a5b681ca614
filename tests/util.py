"""Shared helpers for the parity tests (seeded states, oracle/engine state transfer)."""

from __future__ import annotations

import numpy as np


def _quat_mul(a, b):
  w = a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3]
  x = a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2]
  y = a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1]
  z = a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0]
  return np.stack([w, x, y, z], axis=-1)


def make_states(model, n: int, seed: int = 0, key: str = "robot/init_state", z_range=(-0.06, 0.03),
                tilt: float = 0.15, joint_noise: float = 0.25, vel: float = 0.5):
  """Seeded batch of states around the init keyframe: some envs penetrate the ground, joints are
  perturbed (a few hit their limits), velocities are non-zero, controls are perturbed."""
  rng = np.random.default_rng(seed)
  k = model.keys[key]
  nq, nv, nu = int(model.nq), int(model.nv), int(model.nu)
  qpos = np.tile(k["qpos"], (n, 1))
  qpos[:, 0:2] += rng.uniform(-0.5, 0.5, (n, 2))
  qpos[:, 2] += rng.uniform(*z_range, n)
  ax = rng.normal(size=(n, 3))
  ax /= np.linalg.norm(ax, axis=1, keepdims=True)
  ang = rng.uniform(-tilt, tilt, n)
  dq = np.concatenate([np.cos(ang / 2)[:, None], ax * np.sin(ang / 2)[:, None]], axis=1)
  qpos[:, 3:7] = _quat_mul(qpos[:, 3:7], dq)
  qpos[:, 7:] += rng.uniform(-joint_noise, joint_noise, (n, nq - 7))
  # clamp hinge joints slightly beyond their range for a subset so limit rows get exercised
  lo = model.jnt_range[1:, 0] - 0.02
  hi = model.jnt_range[1:, 1] + 0.02
  qpos[:, 7:] = np.clip(qpos[:, 7:], lo, hi)
  qvel = rng.uniform(-vel, vel, (n, nv))
  ctrl = np.tile(k["ctrl"], (n, 1)) + rng.uniform(-0.3, 0.3, (n, nu))
  warm = rng.uniform(-1.0, 1.0, (n, nv))
  return dict(qpos=qpos, qvel=qvel, ctrl=ctrl, qacc_warmstart=warm)


def terrain_states(model, n: int, seed: int, spread: float):
  """make_states() moved onto random sub-terrain origins of a box-terrain scene (xy spread around them)."""
  rng = np.random.default_rng(seed)
  st = make_states(model, n, seed=seed, z_range=(-0.05, 0.04))
  org = np.asarray(model.arrays["terrain_origins"]).reshape(-1, 3)
  spot = org[rng.integers(0, len(org), n)]
  st["qpos"][:, 0:2] = spot[:, 0:2] + rng.uniform(-spread, spread, (n, 2))
  st["qpos"][:, 2] += spot[:, 2]
  return st


def load_oracle(o, st):
  for k, v in st.items():
    o.field(k)[:] = v


def load_sim(sim, st):
  import torch

  for k, v in st.items():
    getattr(sim.data, k)[:] = torch.as_tensor(v, dtype=torch.float32, device=sim.device)


def relerr(a, b, floor=1e-9):
  """True norm-wise relative error per env: max|a-b| / max|b| (``floor`` only guards an all-zero reference)."""
  a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
  b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
  return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), floor)
