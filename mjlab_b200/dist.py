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
  """Rank-0 logging of the per-env ``(reward, terminated, truncated)`` triple: the one collective of the
  data-parallel path, batched.  Every env step appends one packed ``[n, 3]`` float32 row (the env writes it:
  ``b2_velenv_post`` / ``TrackingFlatEnv`` fill ``env.log_row``) to a device ring of ``every`` rows; one
  ``all_gather_into_tensor`` per ``every`` steps moves the whole ring -> ``out[world, every, n, 3]``.

  Why batched: the physics kernels hold every register of every SM, so a per-step NCCL kernel can neither overlap
  them nor avoid making each env step the max over ranks (r01: +0.21..0.26 ms per env step, efficiency 0.90 at 8
  GPUs).  Logging needs no per-step latency, so the ranks rendezvous once per ``every`` steps (16 by default: the
  gather then costs < 2 % of the r01 figure per step) and the same packing runs at world size 1, which keeps the
  1-GPU baseline comparable.  ``join()`` flushes a partially filled ring (gloo on CPU in the tests, NCCL over
  NVLink on GPUs)."""

  def __init__(self, num_envs: int, device, group=None, every: int = 16, overlap: bool = False):
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.group = group
    self.n = num_envs
    self.every = max(int(every), 1)
    dev = torch.device(device)
    self.ring = torch.zeros((self.every, num_envs, 3), dtype=torch.float32, device=dev)
    self.out = torch.zeros((self.world, self.every, num_envs, 3), dtype=torch.float32, device=dev)
    self.k = 0
    self.flushes = 0
    self._pending = 0  # rows appended since the last gather

  def __call__(self, reward: torch.Tensor, terminated: torch.Tensor | None = None, truncated: torch.Tensor | None = None):
    slot = self.ring[self.k % self.every]
    if terminated is None:
      slot.copy_(reward)  # already packed [n, 3]
    else:
      slot[:, 0] = reward
      slot[:, 1] = terminated.to(torch.float32)
      slot[:, 2] = truncated.to(torch.float32)
    self.k += 1
    self._pending += 1
    if self.k % self.every == 0:
      self.flush()
    return self.out

  def flush(self) -> torch.Tensor:
    if self.world == 1:
      self.out[0].copy_(self.ring)
    else:
      dist.all_gather_into_tensor(self.out.view(-1), self.ring.view(-1), group=self.group)
    self.flushes += 1
    self._pending = 0
    return self.out

  def join(self):
    """Flush rows appended since the last gather (no-op when the ring was just gathered)."""
    if self._pending:
      self.flush()
    return self.out

  @staticmethod
  def summarize(out: torch.Tensor) -> dict:
    return {
      "mean_reward": float(out[..., 0].mean()),
      "terminated": int(out[..., 1].sum()),
      "truncated": int(out[..., 2].sum()),
    }
