"""Developer script (not a test): print per-field CUDA-vs-oracle errors for G1 / Go1."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from mjlab_b200.asset_zoo import load_compiled
from mjlab_b200.sim import Simulation, SimulationCfg
from oracle.oracle import Oracle
sys.path.insert(0, "tests")
from util import load_oracle, load_sim, make_states, relerr

name = sys.argv[1] if len(sys.argv) > 1 else "g1_flat"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = load_compiled(name)
sim = Simulation(n, SimulationCfg(), m, "cuda:0")
sim.set_option("debug_outputs", 1)
o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
print("smem/env bytes", sim.get_option("smem_bytes_per_env"))
st = make_states(m, n, seed=1)
load_oracle(o, st); load_sim(sim, st)
o.forward(); sim.forward(); torch.cuda.synchronize()
d = sim.data
nv = int(m.nv)
def T(x): return x[:].detach().cpu().numpy()
print("ncon  oracle", o.ncon.ravel()[:16], "\n      cuda  ", T(d.ncon)[:16])
print("nefc  oracle", o.nefc.ravel()[:16], "\n      cuda  ", T(d.nefc)[:16])
print("niter oracle", o.solver_niter.ravel()[:16], "\n      cuda  ", T(d.solver_niter)[:16])
for f in ["xpos", "xquat", "xmat", "xipos", "subtree_com", "cvel", "geom_xpos", "geom_xmat", "site_xpos",
          "actuator_force", "qfrc_bias", "qfrc_smooth", "qacc_smooth", "qM", "qfrc_constraint", "qacc",
          "sensordata"]:
  a = T(getattr(d, f)).reshape(n, -1); b = o.field(f).reshape(n, -1)
  e = relerr(a, b)
  print(f"{f:18s} max relerr {e.max():.3e}  median {np.median(e):.3e}  worst env {e.argmax()}")
same = (o.ncon.ravel() == T(d.ncon).ravel())
print("ncon equal in", same.mean())
w = int(np.argmax(relerr(T(d.qacc), o.qacc)))
print("worst env", w, "qacc cuda", T(d.qacc)[w][:8], "\n oracle", o.qacc[w][:8])
print("cost cuda", T(d.solver_cost)[w], "oracle", o.solver_cost[w])
# one step
load_oracle(o, st); load_sim(sim, st)
o.step(); sim.step(); torch.cuda.synchronize()
for f in ["qpos", "qvel", "qacc_warmstart"]:
  e = relerr(T(getattr(d, f)), o.field(f))
  print(f"step {f:14s} max relerr {e.max():.3e} median {np.median(e):.3e}")
for k in range(20):
  o.step(); sim.step()
torch.cuda.synchronize()
for f in ["qpos", "qvel"]:
  e = relerr(T(getattr(d, f)), o.field(f))
  print(f"21 steps {f:10s} max relerr {e.max():.3e} median {np.median(e):.3e}")
print("stats", {k: getattr(sim.stats(), k) for k, _ in sim.stats()._fields_})
