"""Dump the environments with the largest CUDA-vs-oracle error (inputs + both outputs) to
gpurun_out/parity_worst_<cfg>.npz so that they can be replayed in the host emulation of the kernel.
  python tools/parity_worst.py [n_envs]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
from parity_table import WORKLOADS, rel
from util import load_oracle, load_sim, make_states, terrain_states
from mjlab_b200.asset_zoo import load_compiled
from mjlab_b200.sim import Simulation, SimulationCfg
from oracle.oracle import Oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = lambda x: x[:].detach().cpu().numpy()
for cfg, name, kw in WORKLOADS:
  m = load_compiled(name)
  sim = Simulation(n, SimulationCfg(), m, "cuda:0")
  sim.set_option("debug_outputs", 1)
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  kw = dict(kw)
  st = terrain_states(m, n, kw["seed"], kw["spread"]) if "spread" in kw else make_states(m, n, **kw)
  out = {}
  load_oracle(o, st); load_sim(sim, st)
  o.forward(); sim.forward(); torch.cuda.synchronize()
  d = sim.data
  e_acc = np.abs(T(d.qacc) - o.qacc).max(1) / np.abs(o.qacc).max(1)
  fw = dict(g_qacc=T(d.qacc), o_qacc=o.qacc.copy(), g_niter=T(d.solver_niter).ravel(), o_ncon=o.ncon.ravel().copy(),
            g_ncon=T(d.ncon).ravel(), o_nefc=o.nefc.ravel().copy(), g_cost=T(d.solver_cost).ravel(),
            g_qfrc_c=T(d.qfrc_constraint), o_qfrc_c=o.qfrc_constraint.copy())
  load_oracle(o, st); load_sim(sim, st)
  o.step(); sim.step(); torch.cuda.synchronize()
  e_vel = np.abs(T(d.qvel) - o.qvel).max(1) / np.abs(o.qvel).max(1)
  worst = np.unique(np.concatenate([np.argsort(e_acc)[-6:], np.argsort(e_vel)[-6:]]))
  print(cfg, "worst envs", worst.tolist())
  for w in worst:
    print(f"  env {w}: qacc err {e_acc[w]:.2e} qvel err {e_vel[w]:.2e} niter {fw['g_niter'][w]} ncon {fw['g_ncon'][w]}/{fw['o_ncon'][w]} nefc {fw['o_nefc'][w]}")
  np.savez(ROOT / "gpurun_out" / f"parity_worst_{cfg}.npz", worst=worst, e_acc=e_acc, e_vel=e_vel,
           g_qvel=T(d.qvel)[worst], o_qvel=o.qvel[worst], **{f"in_{k}": v[worst] for k, v in st.items()},
           **{k: v[worst] for k, v in fw.items()})
  sim.close()
