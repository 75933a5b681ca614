"""A/B timing of engine options on the bench workload (one process, same states for every variant):
  python tools/time_options.py [n_envs] [preroll_env_steps]
Each line: option set -> us per physics sub-step of step_n(4) (CUDA events, 12 env steps, state restored)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mjlab_b200.envs import VelocityEnvCfg, VelocityFlatEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 100
import os
_wl = os.environ.get("B2_WORKLOAD", "B")  # B: G1 flat, E: Go1 rough boxes, F: Go1 rough boxes + height fields
_kw = dict(B={}, E=dict(robot="go1", terrain="rough"), F=dict(robot="go1", terrain="rough_hf"))[_wl]
env = VelocityFlatEnv(VelocityEnvCfg(num_envs=n, **_kw), device="cuda:0")
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
for _ in range(pre):
  env.step(torch.rand((n, 29 if _wl == "B" else 12), generator=g, device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
sim = env.sim
print("resident_ctas", sim.get_option("resident_ctas"), "smem/env", sim.get_option("smem_bytes_per_env"))
state = {k: getattr(sim.data, k)[:].clone() for k in ("qpos", "qvel", "qacc_warmstart", "ctrl")}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")

def run(label, opts, k=12):
  for key, v in opts.items():
    sim.set_option(key, v)
  for key, v in state.items():
    getattr(sim.data, key)[:] = v
  sim.step_n(4)  # warm
  for key, v in state.items():
    getattr(sim.data, key)[:] = v
  tot = 0.0
  for _ in range(k):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sim.step_n(4); b.record(); torch.cuda.synchronize()
    tot += a.elapsed_time(b)
  st = sim.stats()
  nls = sim.data.solver_nls[:].float().mean().item() if hasattr(sim.data, "solver_nls") else float("nan")
  print(f"{label:44s} {tot / k / 4 * 1e3:8.1f} us/sub-step   ncon {st.ncon_mean:.1f} iters {st.niter_mean:.2f}  ls evals {nls:.2f}")

base = dict(full_solver=0, work_queue=0, split_streams=2, phase_sync=2, reorder_every_substep=1)
run("default: 2 streams", base)
run("1 stream", dict(base, split_streams=1))
run("3 streams", dict(base, split_streams=3))
run("4 streams", dict(base, split_streams=4))
run("2 streams, psync off", dict(base, phase_sync=0))
run("2 streams, full solver", dict(base, full_solver=1))
run("2 streams, ls_relstep off", dict(base, ls_relstep=0))
run("reorder once per step_n", dict(base, ls_relstep=1, reorder_every_substep=0))
run("psync level 1", dict(base, ls_relstep=1, phase_sync=1))
run("psync level 3", dict(base, ls_relstep=1, phase_sync=3))
run("default again", dict(base, ls_relstep=1))
