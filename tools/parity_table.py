"""Per-field error table of the CUDA path against the fp64 oracle.

For each workload (B: g1_flat, C: g1_tracking_flat, E: go1_rough, F: go1_rough_hf) and each build of libb2sim (default +
mjlab_b200/csrc/variants/v_*.so) one forward and one step from identical seeded states on >= 1024 envs;
the error of a field is the norm-wise relative error per env, max|a-b| / max|b| (no floor), reported as
p50 / p99 / max over envs.  Writes gpurun_out/parity_table.json and a markdown table on stdout.

  python tools/parity_table.py [n_envs]            # all builds, one subprocess per build
  python tools/parity_table.py --one <tag> [n]     # (internal) one build in this process
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = [("B", "g1_flat", dict(seed=31)),
             ("C", "g1_tracking_flat", dict(seed=32, tilt=0.5, joint_noise=0.6, vel=1.5)),
             ("E", "go1_rough", dict(seed=33, spread=2.2)),
             ("F", "go1_rough_hf", dict(seed=34, hf_spread=2.2))]
FWD = ["qacc_smooth", "qacc", "qfrc_constraint", "contact_force", "cvel", "qM"]
STEP = ["qpos", "qvel", "qacc_warmstart"]


def rel(a, b, floor=1e-9):
  """Norm-wise relative error per env; envs whose reference is below `floor` have no meaningful relative error and
  are left out (contact forces: contacts that barely touch: a 1e-6 m difference in penetration is 0.05 N; floor 1 N, ~1 % of the robot weight)."""
  import numpy as np
  a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
  b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
  den = np.abs(b).max(axis=1)
  ok = den > floor
  return (np.abs(a - b).max(axis=1)[ok] / den[ok])


def one(tag: str, n: int):
  import numpy as np
  import torch
  from util import hfield_states, load_oracle, load_sim, make_states, terrain_states
  from mjlab_b200.asset_zoo import load_compiled
  from mjlab_b200.sim import Simulation, SimulationCfg
  from oracle.oracle import Oracle

  rows = []
  for cfg, name, kw in WORKLOADS:
    m = load_compiled(name)
    sim = Simulation(n, SimulationCfg(), m, "cuda:0")
    sim.set_option("debug_outputs", 1)
    o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
    # the oracle is run to convergence: with MuJoCo's cap of 10 Newton iterations a few stiff environments stop
    # early in fp64 and their (path-dependent) answer is further from the minimiser than the fp32 engine's
    o.set_option("iterations", 50)
    sim.set_option("iterations", 50)  # (both sides: a capped, unconverged answer depends on the path taken)
    kw = dict(kw)
    if "hf_spread" in kw:
      st = hfield_states(m, n, kw["seed"], kw["hf_spread"])
    elif "spread" in kw:
      st = terrain_states(m, n, kw["seed"], kw["spread"])
    else:
      st = make_states(m, n, **kw)
    load_oracle(o, st)
    load_sim(sim, st)
    o.forward(); sim.forward(); torch.cuda.synchronize()
    T = lambda x: x[:].detach().cpu().numpy()
    d = sim.data
    same = (T(d.ncon).ravel() == o.ncon.ravel())
    # contact sets must be the same contacts: a sphere centre on the mid-plane of a thin box (or a capsule
    # exactly parallel to a face) is pushed out through either face depending on the last bit - such ties are
    # geometry degeneracies of the random test states, not solver error, and are counted separately
    mc = T(d.contact_frame).shape[1]
    fr_o = o.contact_frame.reshape(n, -1, 9)[:, :mc, :3]
    fr_g = T(d.contact_frame).reshape(n, mc, 9)[:, :, :3]
    live = np.arange(mc)[None, :] < o.ncon.reshape(n, 1)
    tie = ((np.abs(fr_g - fr_o).max(-1) > 1e-3) & live).any(1)
    same &= ~tie
    stats = dict(mean_ncon=float(o.ncon.mean()), max_ncon=int(o.ncon.max()), same_ncon=int(same.sum()),
                 geometry_ties=int(tie.sum()), mean_niter=float(T(d.solver_niter).mean()))
    for f in FWD:
      a_f, b_f = T(getattr(d, f)).reshape(n, -1).copy(), o.field(f).reshape(n, -1).copy()
      if f == "contact_force":  # rows beyond ncon are not written by either side (stale values of earlier steps)
        dead = np.repeat(~live, 3, axis=1)
        a_f[:, : dead.shape[1]][dead] = 0.0
        b_f[:, : dead.shape[1]][dead] = 0.0
      e = rel(a_f[same], b_f[same],
              floor=1.0 if f in ("contact_force", "qfrc_constraint") else 1e-9)
      rows.append(dict(build=tag, cfg=cfg, model=name, phase="forward", field=f, n=int(len(e)),
                       p50=float(np.percentile(e, 50)), p99=float(np.percentile(e, 99)), max=float(e.max()), **stats))
    load_oracle(o, st); load_sim(sim, st)
    o.step(); sim.step(); torch.cuda.synchronize()
    for f in STEP:
      e = rel(T(getattr(d, f))[same], o.field(f)[same])
      rows.append(dict(build=tag, cfg=cfg, model=name, phase="step", field=f, n=int(len(e)),
                       p50=float(np.percentile(e, 50)), p99=float(np.percentile(e, 99)), max=float(e.max()), **stats))
    sim.close()
  print("ROWS " + json.dumps(rows))


def main():
  if len(sys.argv) > 2 and sys.argv[1] == "--one":
    return one(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1024)
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  vdir = ROOT / "mjlab_b200" / "csrc" / "variants"
  builds = [("default", None)] + [(p.stem, p) for p in sorted(vdir.glob("v_*.so"))]
  allrows = []
  for tag, path in builds:
    env = dict(os.environ)
    if path:
      env["B2SIM_LIB"] = str(path)
    r = subprocess.run([sys.executable, __file__, "--one", tag, str(n)], capture_output=True, text=True, env=env)
    got = [l for l in r.stdout.splitlines() if l.startswith("ROWS ")]
    if not got:
      print(f"build {tag} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
      continue
    allrows += json.loads(got[0][5:])
  out = ROOT / "gpurun_out"
  out.mkdir(exist_ok=True)
  (out / "parity_table.json").write_text(json.dumps(allrows, indent=1))
  print("| build | cfg | phase | field | envs | p50 | p99 | max |")
  print("|---|---|---|---|---|---|---|---|")
  for r in allrows:
    print(f"| {r['build']} | {r['cfg']} | {r['phase']} | {r['field']} | {r['n']} | {r['p50']:.2e} | {r['p99']:.2e} | {r['max']:.2e} |")


if __name__ == "__main__":
  main()
