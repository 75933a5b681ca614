"""Statistics of the steady-state bench workload that decide where solver work can be cut:
per-env Newton iterations, contacts, active joint limits, and the highest dof touched by any constraint
(contacts: the chains of both bodies; limits: the joint's dof) -> how small the 'constrained prefix' of the
dof vector is (the solver could run in that leading block only).

  python tools/workload_stats.py [n_envs] [preroll_env_steps] [robot]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from mjlab_b200.envs import VelocityEnvCfg, VelocityFlatEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 100
robot = sys.argv[3] if len(sys.argv) > 3 else "g1"
env = VelocityFlatEnv(VelocityEnvCfg(num_envs=n, robot=robot), device="cuda:0")
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
for _ in range(pre):
  env.step(torch.rand((n, env.nu), generator=g, device="cuda:0") * 2 - 1)
m = env.model
A = m.arrays
parent = np.asarray(A["body_parentid"]); dofadr = np.asarray(A["body_dofadr"]); dofnum = np.asarray(A["body_dofnum"])
nb, nv = int(m.nbody), int(m.nv)
top = np.full(nb, -1)  # highest dof index in the chain of a body
for b in range(1, nb):
  top[b] = max(top[parent[b]], dofadr[b] + dofnum[b] - 1 if dofnum[b] else -1)
gb = np.asarray(A["geom_bodyid"])
jr = np.asarray(A["jnt_range"]).reshape(-1, 2); jl = np.asarray(A["jnt_limited"]); jq = np.asarray(A["jnt_qposadr"])
jd = np.asarray(A["jnt_dofadr"]); jt = np.asarray(A["jnt_type"])
hist_kd, hist_it, hist_nc, hist_nl = np.zeros(nv + 1, int), np.zeros(32, int), np.zeros(128, int), np.zeros(64, int)
sim = env.sim
for rep in range(8):
  env.step(torch.rand((n, env.nu), generator=g, device="cuda:0") * 2 - 1)
  torch.cuda.synchronize()
  d = sim.data
  ncon = d.ncon[:].cpu().numpy().ravel(); nit = d.solver_niter[:].cpu().numpy().ravel()
  cg = d.contact_geom[:].cpu().numpy(); qpos = d.qpos[:].cpu().numpy()
  kd = np.full(n, -1)
  for w in range(n):
    k = ncon[w]
    if k:
      kd[w] = max(kd[w], top[gb[cg[w, :k]]].max())
  nl = np.zeros(n, int)
  for j in range(len(jl)):
    if jl[j] and jt[j] == 3:
      v = qpos[:, jq[j]]
      act = (v < jr[j, 0]) | (v > jr[j, 1])
      nl += act
      kd = np.where(act, np.maximum(kd, jd[j]), kd)
  np.add.at(hist_kd, kd + 1, 1); np.add.at(hist_it, np.minimum(nit, 31), 1)
  np.add.at(hist_nc, np.minimum(ncon, 127), 1); np.add.at(hist_nl, np.minimum(nl, 63), 1)
tot = hist_kd.sum()
print("envs x samples", tot)
print("highest constrained dof (kd+1 = prefix length), share of envs:")
cum = 0
for k in range(nv + 1):
  if hist_kd[k]:
    cum += hist_kd[k]
    print(f"  prefix {k:2d}: {hist_kd[k] / tot:6.3f}  cum {cum / tot:6.3f}")
print("newton iterations:", {i: round(hist_it[i] / tot, 3) for i in range(32) if hist_it[i]})
print("ncon:", {i: round(hist_nc[i] / tot, 3) for i in range(128) if hist_nc[i] / tot > 0.004})
print("active limits:", {i: round(hist_nl[i] / tot, 3) for i in range(64) if hist_nl[i]})
print("mean niter", (hist_it * np.arange(32)).sum() / tot, "mean ncon", (hist_nc * np.arange(128)).sum() / tot)
