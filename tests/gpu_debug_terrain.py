"""Ad-hoc diagnostics: dump contact-list differences between the CUDA path and the oracle (run on a GPU box)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from util import load_oracle, load_sim, make_states  # noqa: E402

from mjlab_b200.asset_zoo import load_compiled  # noqa: E402
from mjlab_b200.compiler import Spec  # noqa: E402
from mjlab_b200.sim import Simulation, SimulationCfg  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from test_terrain_gpu import ON_BOX  # noqa: E402
from util import terrain_states as _terrain_states  # noqa: E402


def T(x):
  return x[:].detach().cpu().numpy()


def dump(m, sim, o, n, tag, limit=4):
  d = sim.data
  nc_g, nc_o = T(d.ncon).ravel(), o.ncon.ravel()
  bad = np.nonzero(nc_g != nc_o)[0]
  print(f"[{tag}] ncon mismatch worlds {bad.tolist()} of {n}; overflow gpu {T(d.overflow).ravel().sum()} oracle {o.overflow.sum()}")
  cg, cd, cp, cf = T(d.contact_geom), T(d.contact_dist), T(d.contact_pos), T(d.contact_frame)
  for w in bad[:limit]:
    print(f"  world {w}: gpu {nc_g[w]} oracle {nc_o[w]} ovf gpu {T(d.overflow)[w]} oracle {o.overflow[w]}")
    A = {}
    for k in range(nc_g[w]):
      A.setdefault(tuple(cg[w, k].tolist()), []).append((float(cd[w, k]), cp[w, k].round(5).tolist(), cf[w, k].ravel()[:3].round(4).tolist()))
    B = {}
    og = o.contact_geom[w].reshape(-1, 2)
    for k in range(nc_o[w]):
      B.setdefault(tuple(og[k].tolist()), []).append((float(o.contact_dist[w, k]), o.contact_pos[w, 3 * k:3 * k + 3].round(5).tolist(),
                                                      o.contact_frame[w, 9 * k:9 * k + 3].round(4).tolist()))
    for g in sorted(set(A) | set(B)):
      if len(A.get(g, [])) != len(B.get(g, [])):
        print(f"    pair {g} types {m.geom_type[g[0]]},{m.geom_type[g[1]]} size {m.geom_size[g[0]].tolist()} {m.geom_size[g[1]].tolist()}")
        print(f"      gpu    {A.get(g)}")
        print(f"      oracle {B.get(g)}")
        for gg in g:
          print(f"      geom {gg} pos gpu {T(d.geom_xpos)[w, gg].tolist()} oracle {o.geom_xpos[w, 3 * gg:3 * gg + 3].tolist()}")
          print(f"             mat gpu {T(d.geom_xmat)[w, gg].ravel().round(5).tolist()}")
          print(f"             mat orc {o.geom_xmat[w, 9 * gg:9 * gg + 9].round(5).tolist()}")


def main():
  for name, n, spread in [("go1_stairs_small", 128, 1.4), ("go1_rough", 192, 2.2)]:
    m = load_compiled(name)
    sim = Simulation(n, SimulationCfg(), m, "cuda:0")
    o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
    st = _terrain_states(m, n, 21, spread)
    load_oracle(o, st)
    load_sim(sim, st)
    o.forward()
    sim.forward()
    torch.cuda.synchronize()
    dump(m, sim, o, n, name)
    sim.close()
  geom, z, quat = '<geom type="capsule" size="0.05 0.2" mass="2"/>', 0.549, "0.7071068 0 0.7071068 0"
  m = Spec.from_string(ON_BOX.format(geom=geom, z=z, quat=quat)).compile()
  n = 8
  sim = Simulation(n, SimulationCfg(), m, "cuda:0")
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  rng = np.random.default_rng(2)
  qpos = np.tile(m.qpos0, (n, 1))
  qpos[:, 0:2] += rng.uniform(-0.45, 0.45, (n, 2))
  qpos[:, 2] += rng.uniform(-0.01, 0.005, n)
  st = dict(qpos=qpos, qvel=rng.uniform(-0.2, 0.2, (n, 6)))
  load_oracle(o, st)
  load_sim(sim, st)
  for it in range(40):
    o.forward()
    sim.forward()
    torch.cuda.synchronize()
    if (T(sim.data.ncon).ravel() != o.ncon.ravel()).any():
      dump(m, sim, o, n, f"capsule-on-box it {it}")
      for w in np.nonzero(T(sim.data.ncon).ravel() != o.ncon.ravel())[0]:
        print("   qpos", o.qpos[w].tolist())
      break
    o.step()
    sim.step()
    sim.data.qpos[:] = torch.as_tensor(o.qpos, dtype=torch.float32, device="cuda:0")
    sim.data.qvel[:] = torch.as_tensor(o.qvel, dtype=torch.float32, device="cuda:0")
    sim.data.qacc_warmstart[:] = torch.as_tensor(o.qacc_warmstart, dtype=torch.float32, device="cuda:0")


if __name__ == "__main__":
  main()
