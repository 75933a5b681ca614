"""Small run of the fused MDP kernels (velocity and tracking tasks) for compute-sanitizer:
  compute-sanitizer --tool memcheck python tools/sanitize_env.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from mjlab_b200.envs import VelocityEnvCfg, VelocityFlatEnv  # noqa: E402
from mjlab_b200.envs.tracking_env import TrackingEnvCfg, TrackingFlatEnv  # noqa: E402

n = 64
g = torch.Generator(device="cuda:0")
g.manual_seed(0)
env = VelocityFlatEnv(VelocityEnvCfg(num_envs=n, episode_length_s=0.2), device="cuda:0")
for _ in range(12):  # (episodes of 10 steps: resets and masked forwards are exercised)
  env.step(torch.rand((n, env.nu), generator=g, device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
print("velocity env ok", float(env.episode_length_buf.float().mean()))
trk = TrackingFlatEnv(TrackingEnvCfg(num_envs=n), device="cuda:0")
for _ in range(12):
  trk.step(torch.rand((n, trk.nu), generator=g, device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
print("tracking env ok")
