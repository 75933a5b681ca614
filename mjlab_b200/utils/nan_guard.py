"""NaN guard with the behaviour of the reference's ``utils/nan_guard.py:29-158``: when enabled,
keep a short ring buffer of pre-step states for the first few envs and, the first time a step
produces NaN/Inf in ``qpos/qvel/qacc/qacc_warmstart``, dump them to an ``.npz``.  Disabled by
default, in which case ``watch`` is a no-op context manager.
"""

from __future__ import annotations

import os
import tempfile
import time
from contextlib import contextmanager
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class NanGuardCfg:
  enabled: bool = False
  buffer_size: int = 100
  output_dir: str = os.path.join(tempfile.gettempdir(), "mjlab", "nan_dumps")
  max_envs_to_capture: int = 5


class NanGuard:
  def __init__(self, cfg: NanGuardCfg, num_envs: int, model) -> None:
    self.cfg = cfg
    self.enabled = cfg.enabled
    self.num_envs = num_envs
    self.model = model
    self.num_envs_to_capture = min(num_envs, cfg.max_envs_to_capture)
    self.buffer: list[dict] = []
    self.step_counter = 0
    self._dumped = False
    self.last_dump_path: str | None = None

  def capture(self, data) -> None:
    if not self.enabled:
      return
    n = self.num_envs_to_capture
    self.buffer.append(
      {
        "step": self.step_counter,
        "qpos": data.qpos[:n].detach().cpu().numpy().copy(),
        "qvel": data.qvel[:n].detach().cpu().numpy().copy(),
      }
    )
    if len(self.buffer) > self.cfg.buffer_size:
      self.buffer.pop(0)
    self.step_counter += 1

  def check_and_dump(self, data) -> bool:
    if not self.enabled or self._dumped:
      return False
    bad = torch.zeros(self.num_envs, dtype=torch.bool, device=data.qpos.device)
    for name in ("qpos", "qvel", "qacc", "qacc_warmstart"):
      t = getattr(data, name)
      bad |= ~torch.isfinite(t[:]).all(dim=-1)
    if not bool(bad.any()):
      return False
    self._dump(bad.nonzero().flatten().cpu().numpy())
    self._dumped = True
    return True

  def _dump(self, nan_env_ids: np.ndarray) -> None:
    os.makedirs(self.cfg.output_dir, exist_ok=True)
    stamp = time.strftime("%Y%m%d_%H%M%S")
    path = os.path.join(self.cfg.output_dir, f"nan_dump_{stamp}_{os.getpid()}.npz")
    out = {"nan_env_ids": nan_env_ids, "num_envs_captured": self.num_envs_to_capture}
    for item in self.buffer:
      out[f"qpos_step_{item['step']:06d}"] = item["qpos"]
      out[f"qvel_step_{item['step']:06d}"] = item["qvel"]
    np.savez_compressed(path, **out)
    self.last_dump_path = path
    print(f"[NanGuard] NaN/Inf detected in envs {nan_env_ids[:10].tolist()}; dumped {path}")

  @contextmanager
  def watch(self, data):
    if not self.enabled:
      yield
      return
    self.capture(data)
    yield
    self.check_and_dump(data)
