"""bench.py — env-steps/s of the B200-native physics step on BASELINE.json's headline workload.

Workload (BASELINE.json configs[1]): Unitree G1 velocity-tracking on flat
terrain, 4096 envs per GPU, random-action agent (``2*U(0,1)-1``, reference ``scripts/play.py:167-169``),
dt 0.005 s, decimation 4, Newton <=10 iterations / <=20 line-search iterations, implicitfast,
pyramidal cones, foot-friction domain randomisation, resets on fall (>70 deg) or time-out, pushes.

One "step" = one environment step = ctrl write + 4 physics sub-steps (4 launches of the fused step
kernel) + mask-based partial reset + one forward pass + the (torch) MDP glue of
``VelocityFlatEnv.step``.  ``value`` uses device-resident actions; ``e2e`` feeds actions from pinned
host memory and reads reward/done/obs back every step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs E]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, weak scaling)
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOADS = {
  # BASELINE.json configs[1] (the headline), configs[2], configs[4] (per-GPU share): SURVEY.md §8d B / C / E
  "B": dict(desc="Unitree G1 velocity-tracking flat, {envs} envs/GPU, random-action agent, decimation 4, "
                 "dt 0.005, Newton<=10 it / ls<=20, implicitfast, pyramidal, foot-friction DR, resets+pushes",
            model="g1_flat", robot="g1"),
  "C": dict(desc="Unitree G1 motion-tracking flat (synthetic static clip, self-collision sensor num=10), {envs} envs/GPU, "
                 "random-action agent, decimation 4, dt 0.005, Newton<=10 it / ls<=20, implicitfast, pyramidal, "
                 "DR body_ipos/qpos0/foot friction, RSI resets + pushes, policy obs noise + critic group",
            model="g1_tracking_flat", robot="g1"),
  "E": dict(desc="Unitree Go1 velocity-tracking on the rough box terrain (3564 boxes, curriculum spawn levels <= 5), "
                 "{envs} envs/GPU, random-action agent, decimation 4, dt 0.005, Newton<=10 it / ls<=20, implicitfast, "
                 "pyramidal, foot-friction DR, resets+pushes",
            model="go1_rough", robot="go1"),
}
WORKLOAD = WORKLOADS["B"]["desc"]


def make_env(workload: str, envs: int, seed: int, dev: str):
  from mjlab_b200.envs import TrackingEnvCfg, TrackingFlatEnv, VelocityEnvCfg, VelocityFlatEnv

  if workload == "C":
    return TrackingFlatEnv(TrackingEnvCfg(num_envs=envs, seed=seed), device=dev)
  if workload == "E":
    return VelocityFlatEnv(VelocityEnvCfg(robot="go1", terrain="rough", num_envs=envs, seed=seed), device=dev)
  return VelocityFlatEnv(VelocityEnvCfg(num_envs=envs, seed=seed), device=dev)


def _peaks():
  p = ROOT / "MEASURED_PEAKS.json"
  if p.exists():
    return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
  return 6650.0, "fallback"


class ClockSampler:
  """SM clock / throttle-reason sampling during the timed region (B200_PROFILING.md clocks line).

  NVML is polled from a thread every 5 ms (the timed region of the default run is ~150 ms, shorter than the
  start-up time of an `nvidia-smi -lms` child process, which remains the fallback)."""

  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
       "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index: int):
    self.index = index
    self.proc = None
    self.path = None
    self.thread = None
    self.samples = []      # (sm_mhz, reasons bit mask)
    self.smax = None
    self._stop = False

  def _nvml_handle(self):
    import pynvml
    import torch

    pynvml.nvmlInit()
    try:
      uuid = str(torch.cuda.get_device_properties(self.index).uuid)
      if not uuid.startswith("GPU-"):
        uuid = "GPU-" + uuid
      return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
    except Exception:
      return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

  def start(self):
    try:
      nv, h = self._nvml_handle()
      self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
      nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
      get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

      def loop():
        while not self._stop:
          try:
            self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
          except Exception:
            pass
          time.sleep(0.005)

      self.thread = threading.Thread(target=loop, daemon=True)
      self.thread.start()
      return
    except Exception:
      self.thread = None
    try:
      self.path = tempfile.mktemp(suffix=".csv")
      self.proc = subprocess.Popen(
        ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
         "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
      self.proc = None

  def stop(self):
    if self.thread is not None:
      self._stop = True
      self.thread.join(timeout=2)
      if not self.samples:
        return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": ["no samples"]}
      sm = sorted(x[0] for x in self.samples)
      bits = 0
      for _, r in self.samples:
        bits |= r
      # nvml.h: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
      names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
      return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.smax,
              "reasons": sorted(v for k, v in names.items() if bits & k), "samples": len(sm), "source": "nvml"}
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
      for line in open(self.path):
        f = [x.strip() for x in line.split(",")]
        if len(f) < 9:
          continue
        sm.append(float(f[1]))
        smax.append(float(f[2]))
        for k, nm in enumerate(names):
          if f[5 + k].lower().startswith("active"):
            reasons.add(nm)
    except Exception:
      pass
    if not sm:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons),
            "samples": len(sm), "source": "nvidia-smi"}


class CpuPort:
  """The CPU port of the hot path (oracle, fp32, OpenMP static over envs) on a bounded sample of the
  same workload: keyframe + reset noise, settled onto the ground, random actions, 4 sub-steps per
  env step.  Test infrastructure used here only as the *measured baseline*, never by the product."""

  def __init__(self, envs_per_thread: int = 32, nthreads: int | None = None, workload: str = "B"):
    import re

    import numpy as np

    from mjlab_b200.asset_zoo import g1, go1, load_compiled
    from oracle.oracle import Oracle

    self.np = np
    wl = WORKLOADS[workload]
    zoo = g1 if wl["robot"] == "g1" else go1
    m = load_compiled(wl["model"])
    # all host cores this process may use; torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # otherwise shrink the CPU arm to one thread when the driver launches it under torchrun (N > 1)
    try:
      avail = len(os.sched_getaffinity(0))
    except AttributeError:
      avail = os.cpu_count() or 1
    self.cores = nthreads or avail
    self.n = max(self.cores * envs_per_thread, 8)
    self.o = Oracle(m, nworld=self.n, maxcon=48, precision="f32")
    self.rng = np.random.default_rng(42)
    self.key = m.keys["robot/init_state"]
    qpos = np.tile(self.key["qpos"], (self.n, 1))
    qpos[:, 0:2] += self.rng.uniform(-0.5, 0.5, (self.n, 2))
    if "terrain_origins" in m.arrays:  # rough terrain: spawn on sub-terrain origins of the lower levels
      org = np.asarray(m.arrays["terrain_origins"])[:6].reshape(-1, 3)
      qpos[:, 0:3] += org[self.rng.integers(0, len(org), self.n)]
    yaw = self.rng.uniform(-3.14, 3.14, self.n)
    qpos[:, 3], qpos[:, 6] = np.cos(yaw / 2), np.sin(yaw / 2)
    self.o.qpos[:] = qpos
    names = [x.split("/")[-1] for x in m.names["joint"][1:]]
    self.scale = np.array(
      [next((v for p, v in zoo.ACTION_SCALE.items() if re.match(p, nm)), 0.5) for nm in names])
    self.o.ctrl[:] = self.key["ctrl"]
    for _ in range(40):  # settle onto the ground (untimed) so the sample carries the contact load
      self.o.step(self.cores)

  def env_step(self) -> float:
    t0 = time.perf_counter()
    self.o.ctrl[:] = self.key["ctrl"] + self.scale * self.rng.uniform(-1, 1, (self.n, len(self.scale)))
    for _ in range(4):
      self.o.step(self.cores)
    return time.perf_counter() - t0

  def describe(self, env_steps: int, seconds: float) -> dict:
    return {
      "value": self.n * env_steps / seconds, "unit": UNIT, "cores": self.cores, "kind": "port",
      "sample": f"{self.n} envs x {env_steps} env-steps (x4 sub-steps), fp32 restated CPU oracle "
                f"(not C-MuJoCo, not mujoco_warp-CPU), OpenMP static over envs, {seconds:.1f} s",
    }


def cpu_reference_throughput(env_steps: int = 12, workload: str = "B"):
  port = CpuPort(workload=workload)
  port.env_step()  # warm caches / thread pool
  dt = sum(port.env_step() for _ in range(env_steps))
  return port.describe(env_steps, dt)


def run_reference(args):
  """--impl reference: the reference's physics cannot be installed here (mujoco / mujoco_warp /
  warp wheels absent, no network), so this arm times the CPU port of the same path on
  all host cores; each step is one env step of a bounded sample of the workload.  Rank 0 only."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  t0 = time.perf_counter()
  port = CpuPort(workload=args.workload)
  for _ in range(max(args.warmup, 1)):
    port.env_step()
  K = min(args.steps, 40)  # bounded: the whole arm must finish within a few minutes
  dt = sum(port.env_step() for _ in range(K))
  cb = port.describe(K, dt)
  v = cb["value"]
  line = {
    "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
    "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * dt / K,
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
    "data": "synthetic", "config": {"workload": WORKLOADS[args.workload]["desc"].format(envs=args.envs),
                                     "note": "CPU port on host cores; each step = one env step of a "
                                             f"bounded sample ({port.n} envs)"},
    "cpu_baseline": cb,
    "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "wall_s": time.perf_counter() - t0,
  }
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=10)
  ap.add_argument("--impl", default="b200")
  ap.add_argument("--envs", type=int, default=4096)
  ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS), help="SURVEY.md §8d config: B (headline), C, E")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-flush", action="store_true")
  ap.add_argument("--preroll", type=int, default=100, help="untimed env steps before warm-up")
  ap.add_argument("--no-graph", action="store_true", help="do not capture the env step in a CUDA graph")
  args = ap.parse_args()
  if args.impl == "reference":
    return run_reference(args)

  import torch
  import torch.distributed as dist

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = f"cuda:{local}"
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(dev))
  W = max(args.warmup, 3)
  K = args.steps
  env = make_env(args.workload, args.envs, 42 + rank, dev)
  n, nu = env.num_envs, env.nu
  gen = torch.Generator(device=dev)
  gen.manual_seed(1234 + rank)
  flush_buf = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  from mjlab_b200.dist import EnvLogGather

  # the one collective of the data-parallel path: per-env (reward, done) to rank 0 for logging
  gathers = {"device": EnvLogGather(n, dev, every=16), "host": EnvLogGather(n, dev, every=16)}

  def run(kind: str, steps: int, timed: bool):
    """kind: 'device' (actions resident) or 'host' (pinned host actions, results read back)."""
    ev = []
    phys_ms = 0.0
    host_actions = None
    # stream-ordered when steps are enqueued back to back, side stream when the host syncs every step
    gather_logs = gathers[kind]
    if kind == "host":
      host_actions = [(torch.rand((n, nu)) * 2 - 1).pin_memory() for _ in range(steps)]
      out_r = torch.empty(n, pin_memory=True)
      out_d = torch.empty((n, 2), dtype=torch.bool).pin_memory()
      out_o = None
    for k in range(steps):
      if flush_buf is not None:
        flush_buf.zero_()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      if kind == "host":
        action = host_actions[k].to(dev, non_blocking=True)
      else:
        action = torch.rand((n, nu), generator=gen, device=dev) * 2 - 1
      obs, reward, terminated, truncated, _ = env.step(action)
      gather_logs(env.log_row)  # packed (reward, terminated, truncated), written by the env step itself
      if kind == "host":
        out_r.copy_(reward, non_blocking=True)
        out_d.copy_(torch.stack([terminated, truncated], dim=1), non_blocking=True)
        if out_o is None:
          out_o = torch.empty(obs.shape, pin_memory=True)
        out_o.copy_(obs, non_blocking=True)
        e1.record()
        e1.synchronize()  # the caller consumes the result before issuing the next action
      else:
        e1.record()
      ev.append((e0, e1))
    # the log gather runs on a side stream; its tail after the last step belongs to the timed region
    j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    j0.record()
    gather_logs.join()
    j1.record()
    torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in ev) + j0.elapsed_time(j1)
    return total, (n * nu * 4, n * 4 + n * 2 + (obs.numel() * 4)) if kind == "host" else None

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def maxr(x: float) -> float:
    if world == 1:
      return x
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  # launches of OUR kernels per env step (counted eagerly; graph replays do not pass through Python)
  l0 = env.sim.launch_count()
  run("device", 1, False)
  launches_per_step = env.sim.launch_count() - l0

  # ---- physics-kernel timing hook: events around the 4 sub-step launches ---------------------------
  phys_events = []
  orig_step_n = env.sim.step_n

  def timed_step_n(k):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig_step_n(k)
    b.record()
    phys_events.append((a, b, k))

  # Pre-roll to the steady-state mix of standing / falling / freshly reset robots (all envs start
  # standing at t=0, so without it the timed window would measure a transient), then W warm-up steps.
  run("device", args.preroll, False)
  # kernel-only timing pass (eager, events around the sub-step launches): feeds the roofline
  env.sim.step_n = timed_step_n
  run("device", max(W, 5), False)
  env.sim.step_n = orig_step_n
  kern_ms = sum(a.elapsed_time(b) for a, b, _ in phys_events) / max(sum(k for _, _, k in phys_events), 1)
  kern_ms = maxr(kern_ms)
  if not args.no_graph:
    env.enable_cuda_graph()
  run("device", W, False)  # warm-up (untimed)
  sampler = ClockSampler(local)
  barrier()
  if rank == 0:
    sampler.start()
  wall0 = time.perf_counter()
  total_ms, _ = run("device", K, True)
  barrier()
  wall = time.perf_counter() - wall0
  clocks = sampler.stop() if rank == 0 else None
  launches = launches_per_step * K
  total_ms = maxr(total_ms)
  st = env.sim.stats()
  import ctypes
  sb, wb = ctypes.c_double(), ctypes.c_double()
  env.sim._lib.b2_algorithmic_bytes(env.sim._h, env.sim._stream(), ctypes.byref(sb), ctypes.byref(wb))

  # ---- end-to-end through host buffers -----------------------------------------------------------
  run("host", W, False)
  barrier()
  e2e_ms, (h2d, d2h) = run("host", K, True)
  barrier()
  e2e_ms = maxr(e2e_ms)

  if rank == 0:
    peak, which = _peaks()
    value = world * n * K / (total_ms * 1e-3)
    e2e = world * n * K / (e2e_ms * 1e-3)
    achieved = wb.value / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "step_kernel_traffic.json"
    if tp.exists():
      traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
    line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
      "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32", "data": "synthetic",
      "config": {
        "workload": WORKLOADS[args.workload]["desc"].format(envs=n), "config_id": args.workload,
        "envs_per_gpu": n, "decimation": 4,
        "parallelism": f"dp{world} (envs sharded, no physics coupling; (reward, done) rows all-gathered to every rank "
                       "once per 16 env steps, same packing at 1 GPU)",
        "env_step": "one CUDA-graph replay per env step" if not args.no_graph else "eager",
        "l2": "flushed between timed steps (256 MiB memset, untimed)" if flush_buf is not None else "not flushed",
        "preroll_env_steps": args.preroll,
        "mean_ncon": st.ncon_mean, "mean_nefc": st.nefc_mean, "mean_newton_iters": st.niter_mean,
        "overflow_worlds": st.overflow_worlds,
      },
      "clocks": clocks,
      "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
              "ms_per_step": e2e_ms / K},
      "gpu_launches": int(launches),
      "roofline": {
        "kernel": "b2_step_kernel<true> (fused physics sub-step)", "bound": "hbm",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": which, "traffic": traffic, "kernel_ms": kern_ms,
        "algorithmic_bytes_per_launch": wb.value,
        "note": "whole-step minimal traffic (state in/out + consumer-visible kinematics); the kernel is "
                "latency/ALU bound, the constraint Jacobian never exists in HBM",
        "solver_formula_gbs": sb.value / (kern_ms * 1e-3) / 1e9,
      },
      "physics_only_env_steps_per_sec": world * n / (4 * kern_ms * 1e-3),
      "wall_s": wall,
    }
    if not args.no_cpu_baseline and world == 1:
      try:
        line["cpu_baseline"] = cpu_reference_throughput(workload=args.workload)
      except Exception as e:  # the oracle is test infrastructure; never fail the bench on it
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                "sample": f"failed: {e}"}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
