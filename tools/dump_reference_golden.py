"""Pin the oracle (and the CUDA path) against the REAL reference engine — to be run on any machine that has
``mujoco`` (C MuJoCo, fp64) and, optionally, ``mujoco_warp`` + the reference ``mjlab`` package.  None of them
exists in the authoring image, which is why DESIGN.md says "parity unpinned"; this is the one
command that changes that:

    python tools/dump_reference_golden.py            # writes tests/golden_ref/<scene>_seed<k>.npz
    python -m pytest tests/test_golden_ref.py        # oracle (CPU) and CUDA path (-m gpu) against those files

What it does, per scene (g1_flat, g1_tracking_flat, go1_flat):
  1. builds the scene with the reference's own code path — ``mjlab.scene.Scene`` + the zoo's robot cfg +
     ``MujocoCfg.edit_spec`` exactly as ``ManagerBasedEnv.__init__`` does (``scene/scene.py:24-40``,
     ``sim/sim.py:65-82``, ``tasks/velocity/velocity_env_cfg.py:248-256``) — and compiles it with C MuJoCo;
  2. compares the compiled ``mjModel`` with this repo's compiled blob (``mjlab_b200/asset_zoo/compiled``): sizes,
     element names/order, masses, ``dof_invweight0`` / ``body_invweight0`` / ``stat.meaninertia`` (the compiler
     constants that scale the contact regulariser) — printed as max abs / rel differences and stored;
  3. loads the seeded states of ``tests/util.make_states`` (the states every parity test uses) and records
     ``mj_forward`` + ``mj_step`` (call sites mirrored: ``sim/sim.py:106-107``) per env, in fp64;
  4. when ``mujoco_warp`` imports: the same through ``mjwarp.put_model / put_data / forward / step``
     (``sim/sim.py:110-119,136,139``), nworld = n, fp32 — the path north_star's 1e-4 is stated against.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
OUT = ROOT / "tests" / "golden_ref"

SCENES = {
  # blob name -> (robot cfg import path, task sim options, extra sensors)
  "g1_flat": ("mjlab.asset_zoo.robots.unitree_g1.g1_constants", "G1_ROBOT_CFG", "velocity"),
  "g1_tracking_flat": ("mjlab.asset_zoo.robots.unitree_g1.g1_constants", "G1_ROBOT_CFG", "tracking"),
  "go1_flat": ("mjlab.asset_zoo.robots.unitree_go1.go1_constants", "GO1_ROBOT_CFG", "velocity"),
}
FWD = ("qacc", "qacc_smooth", "qfrc_bias", "qfrc_constraint", "actuator_force", "xpos", "xquat", "cvel", "subtree_com",
       "sensordata")
STEP = ("qpos", "qvel", "qacc_warmstart")


def build_reference_model(blob: str):
  """The scene the reference's env would compile for this blob (plane terrain + one robot, task options)."""
  import importlib
  from dataclasses import replace

  import mujoco
  from mjlab.scene import Scene, SceneCfg
  from mjlab.sim import MujocoCfg
  from mjlab.terrains import TerrainImporterCfg
  from mjlab.utils.spec_config import ContactSensorCfg

  mod, attr, task = SCENES[blob]
  cfg = getattr(importlib.import_module(mod), attr)
  if task == "tracking":  # tasks/tracking/config/g1/flat_env_cfg.py:11-20
    cfg = replace(cfg, sensors=(ContactSensorCfg(name="self_collision", subtree1="pelvis", subtree2="pelvis",
                                                  data=("found",), reduce="netforce", num=10),))
  else:  # tasks/velocity/config/*/rough_env_cfg.py: one foot-ground contact sensor per foot
    from mjlab_b200.asset_zoo import g1, go1

    zoo = g1 if "g1" in blob else go1
    cfg = replace(cfg, sensors=tuple(
      ContactSensorCfg(name=s.name, **{k: v for k, v in s.__dict__.items() if k != "name" and v is not None})
      for s in zoo.velocity_sensors()))
  scene = Scene(SceneCfg(terrain=TerrainImporterCfg(terrain_type="plane"), num_envs=1, entities={"robot": cfg}), device="cpu")
  MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20).edit_spec(scene.spec)
  return mujoco, scene.compile()


def compare_models(mjm, mine) -> dict:
  """Differences between C MuJoCo's compiled model and this repo's compiler output."""
  import mujoco

  rep = {}
  for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite"):
    rep[f"size_{k}"] = (int(getattr(mjm, k)), int(getattr(mine, k)))
  names = lambda kind, n: [mujoco.mj_id2name(mjm, kind, i) or "" for i in range(n)]
  rep["body_names_equal"] = names(mujoco.mjtObj.mjOBJ_BODY, mjm.nbody) == list(mine.names["body"])
  rep["joint_names_equal"] = names(mujoco.mjtObj.mjOBJ_JOINT, mjm.njnt) == list(mine.names["joint"])
  for f in ("body_mass", "body_inertia", "body_ipos", "body_subtreemass", "body_invweight0", "dof_invweight0",
            "dof_armature", "jnt_range", "actuator_gainprm", "actuator_biasprm", "geom_size", "geom_friction"):
    a, b = np.asarray(getattr(mjm, f), dtype=np.float64), np.asarray(getattr(mine, f), dtype=np.float64).reshape(np.shape(getattr(mjm, f)))
    rep[f"maxabs_{f}"] = float(np.abs(a - b).max()) if a.size else 0.0
    rep[f"maxrel_{f}"] = float((np.abs(a - b) / np.maximum(np.abs(a), 1e-12)).max()) if a.size else 0.0
  rep["meaninertia"] = (float(mjm.stat.meaninertia), float(mine.stat_meaninertia))
  return rep


def dump(blob: str, n: int = 64, seed: int = 101) -> Path:
  from util import make_states

  from mjlab_b200.asset_zoo import load_compiled

  mujoco, mjm = build_reference_model(blob)
  mine = load_compiled(blob)
  rep = compare_models(mjm, mine)
  for k, v in rep.items():
    print(f"  {blob}: {k} = {v}")
  st = make_states(mine, n, seed=seed)
  out = {f"in_{k}": v for k, v in st.items()}
  d = mujoco.MjData(mjm)
  rec = {f"c_fwd_{f}": [] for f in FWD}
  rec.update({f"c_step_{f}": [] for f in STEP})
  rec.update(c_ncon=[], c_nefc=[], c_niter=[])
  for w in range(n):
    mujoco.mj_resetData(mjm, d)
    d.qpos[:], d.qvel[:], d.ctrl[:], d.qacc_warmstart[:] = st["qpos"][w], st["qvel"][w], st["ctrl"][w], st["qacc_warmstart"][w]
    mujoco.mj_forward(mjm, d)
    for f in FWD:
      rec[f"c_fwd_{f}"].append(np.array(getattr(d, f)).ravel().copy())
    rec["c_ncon"].append(d.ncon); rec["c_nefc"].append(d.nefc); rec["c_niter"].append(int(d.solver_niter[0]))
    d.qpos[:], d.qvel[:], d.ctrl[:], d.qacc_warmstart[:] = st["qpos"][w], st["qvel"][w], st["ctrl"][w], st["qacc_warmstart"][w]
    mujoco.mj_step(mjm, d)
    for f in STEP:
      rec[f"c_step_{f}"].append(np.array(getattr(d, f)).copy())
  out.update({k: np.array(v) for k, v in rec.items()})
  try:
    import mujoco_warp as mjwarp
    import warp as wp

    m = mjwarp.put_model(mjm)
    dd = mjwarp.put_data(mjm, mujoco.MjData(mjm), nworld=n, nconmax=64 * n, njmax=300)

    def load():
      for k in ("qpos", "qvel", "ctrl", "qacc_warmstart"):
        getattr(dd, k).assign(wp.array(st[k].astype(np.float32), dtype=float))

    load(); mjwarp.forward(m, dd); wp.synchronize()
    for f in ("qacc", "qfrc_constraint", "sensordata"):
      out[f"w_fwd_{f}"] = getattr(dd, f).numpy().reshape(n, -1)
    load(); mjwarp.step(m, dd); wp.synchronize()
    for f in STEP:
      out[f"w_step_{f}"] = getattr(dd, f).numpy().reshape(n, -1)
    out["w_version"] = np.array(getattr(mjwarp, "__version__", "unknown"))
  except ImportError:
    print("  (mujoco_warp not importable: C MuJoCo results only)")
  out["model_report"] = np.array(repr(rep))
  out["mujoco_version"] = np.array(mujoco.__version__)
  OUT.mkdir(exist_ok=True)
  path = OUT / f"{blob}_seed{seed}.npz"
  np.savez_compressed(path, **out)
  return path


if __name__ == "__main__":
  try:
    import mujoco  # noqa: F401
    if getattr(mujoco, "__b2_compat__", False):
      raise ImportError("only the compat stand-in is present")
    import mjlab  # noqa: F401
  except ImportError as e:
    sys.exit(f"dump_reference_golden: needs the real `mujoco` and the reference `mjlab` package ({e}); "
             "run it on a machine that has them and commit tests/golden_ref/*.npz")
  for blob in SCENES:
    print("wrote", dump(blob))
