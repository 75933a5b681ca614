"""Multi-GPU harness: environments shard trivially across ranks (one process per GPU, each with
its own ``Simulation``); the only collective on the data path is one all-gather of the per-env
``(reward, terminated, truncated)`` triple so rank 0 can log the whole job (north_star; SURVEY.md
§8e).  The reference has no multi-GPU code at all (§2.2) — this is new harness code outside the
mjlab surface.  Backend: NCCL over NVLink on GPUs, gloo on CPU (tests).
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend: str | None = None, device: torch.device | None = None) -> tuple[int, int]:
  """Initialise the default process group from torchrun's environment. Returns (rank, world)."""
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  if world > 1 and not dist.is_initialized():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
  return rank, world


def shard_range(num_envs_total: int, rank: int, world: int) -> tuple[int, int]:
  """Rank r owns envs [r*N/W, (r+1)*N/W); N must divide evenly (weak scaling uses N = W * n)."""
  if num_envs_total % world:
    raise ValueError(f"num_envs_total={num_envs_total} is not divisible by world={world}")
  per = num_envs_total // world
  return rank * per, (rank + 1) * per


class EnvLogGather:
  """Preallocated all-gather of (reward f32, terminated, truncated) -> ``[world, n, 3]`` float32."""

  def __init__(self, num_envs: int, device, group=None):
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.group = group
    self.n = num_envs
    self.packed = torch.empty((num_envs, 3), dtype=torch.float32, device=device)
    self.out = torch.empty((self.world, num_envs, 3), dtype=torch.float32, device=device)

  def __call__(self, reward: torch.Tensor, terminated: torch.Tensor, truncated: torch.Tensor):
    self.packed[:, 0] = reward
    self.packed[:, 1] = terminated.to(torch.float32)
    self.packed[:, 2] = truncated.to(torch.float32)
    if self.world == 1:
      self.out[0] = self.packed
    else:
      dist.all_gather_into_tensor(self.out.view(self.world * self.n, 3), self.packed, group=self.group)
    return self.out

  @staticmethod
  def summarize(out: torch.Tensor) -> dict:
    return {
      "mean_reward": float(out[..., 0].mean()),
      "terminated": int(out[..., 1].sum()),
      "truncated": int(out[..., 2].sum()),
    }
