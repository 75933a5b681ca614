"""Small run of the step kernel for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool racecheck python tools/sanitize_step.py go1_rough_hf 64
Height-field terrain by default: the lane-distributed narrowphase keeps owner records, a queue and the sort buffer of
the per-field broadphase in shared memory between warp-level synchronisation points."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

from mjlab_b200.asset_zoo import load_compiled  # noqa: E402
from mjlab_b200.sim import Simulation, SimulationCfg  # noqa: E402
from util import hfield_states, load_sim, make_states, terrain_states  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "go1_rough_hf"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = load_compiled(name)
sim = Simulation(n, SimulationCfg(nconmax=48 * n), m, "cuda:0")
if "hf" in name:
  st = hfield_states(m, n, 5, 2.2, clearance=(-0.08, 0.02))
elif "rough" in name:
  st = terrain_states(m, n, 5, 2.2)
else:
  st = make_states(m, n, seed=5)
load_sim(sim, st)
sim.forward()
for _ in range(3):
  sim.step()
torch.cuda.synchronize()
print("ncon", sim.data.ncon[:].flatten().tolist()[:16], "finite", bool(torch.isfinite(sim.data.qpos[:]).all()))
