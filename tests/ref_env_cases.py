"""Run by tests/test_reference_own_tests.py through tests/ref_runner.py (not collected by the main suite: it needs the
stand-ins ref_runner installs).  The reference's ``ManagerBasedRlEnv`` - scene, entity, action / observation / reward /
termination / command / event managers, ``Simulation`` - constructed from its own Go1 flat velocity task config and
stepped on this repo's engine (the CUDA source compiled for the host): the whole hot loop of SURVEY.md §3.2."""
import numpy as np
import torch


import pytest


@pytest.fixture(scope="module")
def env(tmp_path_factory):
  """One env per interpreter: the task configs share their SceneEntityCfg default instances, which a first
  construction resolves in place (a second ManagerBasedRlEnv in the same process fails upstream too)."""
  import os

  from mjlab.envs.manager_based_rl_env import ManagerBasedRlEnv

  if os.environ.get("B2_REF_TASK", "go1") == "g1":  # BASELINE config B: G1 velocity tracking on flat ground
    from mjlab.tasks.velocity.config.g1.flat_env_cfg import UnitreeG1FlatEnvCfg as Cfg
  elif os.environ.get("B2_REF_TASK") == "g1_tracking":  # BASELINE config C: G1 motion tracking (static synthetic clip)
    from mjlab.tasks.tracking.config.g1.flat_env_cfg import G1FlatEnvCfg as Cfg
  elif os.environ.get("B2_REF_TASK") == "go1_rough":  # BASELINE config E: Go1 on the generated rough terrain
    from mjlab.tasks.velocity.config.go1.rough_env_cfg import UnitreeGo1RoughEnvCfg as Cfg
  else:
    from mjlab.tasks.velocity.config.go1.flat_env_cfg import UnitreeGo1FlatEnvCfg as Cfg
  torch.manual_seed(0)
  cfg = Cfg()
  cfg.scene.num_envs = 4
  if not hasattr(cfg.commands, "motion"):
    cfg.episode_length_s = 15 * cfg.decimation * cfg.sim.mujoco.timestep  # 15 env steps: time-outs fall inside the test
  if hasattr(cfg.commands, "motion"):
    cfg.commands.motion.motion_file = _static_clip(tmp_path_factory.mktemp("clip") / "clip.npz")
  e = ManagerBasedRlEnv(cfg, device=os.environ.get("B2_REF_DEVICE", "cpu"))
  yield e
  e.close()


def _static_clip(path, frames: int = 200):
  """A motion file in the layout of tasks/tracking/mdp/commands.py:30-50 holding the robot standing still in its
  initial keyframe (the real clips come from a wandb registry): body poses from this repo's host kinematics."""
  from mjlab_b200.asset_zoo import load_compiled
  from mjlab_b200.compiler.compile import kinematics_qpos0

  m = load_compiled("g1_tracking_flat")
  q = np.array(m.keys["robot/init_state"]["qpos"], dtype=float)
  kin = kinematics_qpos0(m, q)
  first = m.names["body"].index("robot/pelvis")  # robot bodies only (world and the terrain body dropped)
  xpos, xquat = np.asarray(kin[0])[first:], np.asarray(kin[1])[first:]
  nb = len(xpos)
  np.savez(path, fps=np.array([50]), joint_pos=np.tile(q[7:], (frames, 1)), joint_vel=np.zeros((frames, len(q) - 7)),
           body_pos_w=np.tile(xpos, (frames, 1, 1)), body_quat_w=np.tile(xquat, (frames, 1, 1)),
           body_lin_vel_w=np.zeros((frames, nb, 3)), body_ang_vel_w=np.zeros((frames, nb, 3)))
  return str(path)


def test_reference_env_steps_on_the_engine(env):
  assert type(env.sim).__module__ == "mjlab.sim.sim" and type(env.scene).__module__ == "mjlab.scene.scene"
  obs, _ = env.reset()
  nact = env.action_manager.total_action_dim
  assert nact == int(env.sim.mj_model.nu) and env.cfg.decimation == 4
  assert set(obs) == {"policy", "critic"} and all(v.shape[0] == 4 and v.shape[1] >= 3 * nact for v in obs.values())
  if nact == 12:
    assert {k: tuple(v.shape) for k, v in obs.items()} == {"policy": (4, 48), "critic": (4, 48)}
  g = torch.Generator().manual_seed(0)
  dev = env.device
  dt = float(env.sim.mj_model.opt_timestep)
  tracking, resets = hasattr(env.cfg.commands, "motion"), 0
  for k in range(30):
    obs, rew, term, trunc, info = env.step((torch.rand((4, nact), generator=g) * 2 - 1).to(dev))
    assert rew.shape == (4,) and torch.isfinite(rew).all() and all(torch.isfinite(v).all() for v in obs.values())
    if tracking:
      resets += int((term | trunc).sum())  # off the clip by more than the task's thresholds: terminated, reset onto the clip
    else:
      # 0.3 s episodes of random actions from the standing pose: nobody falls; everybody times out at steps 15 and 30,
      # which runs the reset events (root pose / joints), the command resampling and sim.forward() of the reset path
      assert not term.any() and bool(trunc.all()) == (k % 15 == 14) and bool(trunc.any()) == (k % 15 == 14), (k, term, trunc)
  assert abs(float(env.sim.data.time[0]) - 30 * 4 * dt) < 1e-4
  if tracking:
    assert resets > 0 and (env.episode_length_buf < 30).any()  # the termination -> reset (RSI) -> forward path ran
  else:
    assert (env.episode_length_buf == 0).all() and env.max_episode_length == 15


def test_reference_env_physics_matches_the_oracle(env):
  """The states the reference env reaches are the oracle's: same compiled model, the ctrl the action manager wrote,
  four sub-steps per env step (no reset, push or randomisation falls inside the compared window)."""
  from oracle.oracle import Oracle

  env.reset()
  m = env.sim.mj_model
  o = Oracle(m, nworld=4, maxcon=64)
  for f in ("qpos", "qvel", "qacc_warmstart"):
    o.field(f)[:] = getattr(env.sim.data, f)[:].cpu().numpy()
  for name in ("geom_friction", "body_mass", "body_ipos", "dof_armature", "qpos0"):  # per-world model fields, if expanded
    t = getattr(env.sim.model, name)[:]
    if t.shape[0] == 4:
      o.model_field(name)[:] = t.cpu().numpy().reshape(4, -1)
  nact = env.action_manager.total_action_dim
  g = torch.Generator().manual_seed(1)
  compared = 0
  for k in range(8):
    _, _, term, trunc, _ = env.step((torch.rand((4, nact), generator=g) * 0.5 - 0.25).to(env.device))
    o.field("ctrl")[:] = env.sim.data.ctrl[:].cpu().numpy()
    for _ in range(env.cfg.decimation):
      o.step()
    done = (term | trunc).cpu().numpy()
    keep = ~done
    compared += int(keep.sum())
    err = np.abs(env.sim.data.qpos[:].cpu().numpy() - o.qpos)[keep].max() if keep.any() else 0.0
    assert err < 2e-3, (k, err)
    for f in ("qpos", "qvel", "qacc_warmstart"):  # an env that was reset (tracking: off the clip) restarts the oracle there
      o.field(f)[done] = getattr(env.sim.data, f)[:].cpu().numpy()[done]
  assert compared >= 24
