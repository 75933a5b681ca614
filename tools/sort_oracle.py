"""How much would a perfect predictor of solver work buy the heavy-first / CTA-grouping sort?
The order kernel keys on the previous launch's Newton iterations.  Here the same state is stepped twice: the first run
is ordered by the previous (different) step's counts - the normal situation - the second by the counts of this very
state (restored), i.e. a perfect prediction.  Prints both sub-step times.

  python tools/sort_oracle.py [n_envs] [preroll_env_steps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mjlab_b200.envs import VelocityEnvCfg, VelocityFlatEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 100
env = VelocityFlatEnv(VelocityEnvCfg(num_envs=n), device="cuda:0")
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
for _ in range(pre):
  env.step(torch.rand((n, 29), generator=g, device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
sim = env.sim
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
keys = ("qpos", "qvel", "qacc_warmstart", "ctrl")

def timed():
  flush.zero_()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record(); sim.step_n(1); b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) * 1e3

tn = tp = 0.0
K = 24
for k in range(K):
  for _ in range(3):  # move on to a new state (and leave its predecessor's counts in the key fields)
    env.step(torch.rand((n, 29), generator=g, device="cuda:0") * 2 - 1)
  st = {f: getattr(sim.data, f)[:].clone() for f in keys}
  tn += timed()                      # ordered by the previous step's counts
  for f, v in st.items():
    getattr(sim.data, f)[:] = v
  tp += timed()                      # ordered by this state's own counts
print(f"sub-step, order from the previous step: {tn / K:7.1f} us;  order from the step's own counts: {tp / K:7.1f} us")
