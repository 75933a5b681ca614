"""Box primitives + grid-static broadphase: CUDA path vs CPU oracle (BASELINE config E groundwork).

Tolerances: as test_parity_gpu.py near the origin; the full rough terrain spans +-100 m, where fp32 world
coordinates resolve 8e-6 m, so contact distances are compared to 5e-5 m and accelerations to 5e-3 there."""

import numpy as np
import pytest
import torch

from util import load_oracle, load_sim, make_states, relerr
from util import terrain_states as _terrain_states

pytestmark = pytest.mark.gpu


def T(x):
  return x[:].detach().cpu().numpy()


def _contacts(ncon, geom, dist, w):
  k = int(ncon[w])
  return sorted(zip(map(tuple, np.asarray(geom[w]).reshape(-1, 2)[:k].tolist()), np.asarray(dist[w])[:k].tolist()))


def _check_contacts(sim, o, n, dist_tol):
  d = sim.data
  nc = o.ncon.ravel()
  assert (T(d.ncon).ravel() == nc).all()
  assert (T(d.overflow).ravel() == o.overflow.ravel()).all()
  cg, og = T(d.contact_geom), o.contact_geom.reshape(n, -1, 2)
  cd = T(d.contact_dist)
  same_order = 0
  for w in range(n):
    a, b = _contacts(nc, cg, cd, w), _contacts(nc, og, o.contact_dist, w)
    assert [x[0] for x in a] == [x[0] for x in b]
    if nc[w]:
      assert np.abs(np.array([x[1] for x in a]) - np.array([x[1] for x in b])).max() < dist_tol
    same_order += int((cg[w, : nc[w]] == og[w, : nc[w]]).all())
  return same_order


ON_BOX = """
<mujoco><option timestep="0.002"/>
  <worldbody>
    <body name="table" pos="0 0 0.25"><geom name="top" type="box" size="0.5 0.5 0.25"/></body>
    <body name="obj" pos="0 0 {z}" quat="{quat}"><freejoint/>{geom}</body>
  </worldbody></mujoco>
"""


@pytest.mark.parametrize("geom,z,quat", [
  ('<geom type="sphere" size="0.1" mass="2"/>', 0.599, "1 0 0 0"),
  ('<geom type="capsule" size="0.05 0.2" mass="2"/>', 0.549, "0.7071068 0 0.7071068 0"),
  ('<geom type="capsule" size="0.05 0.2" mass="2"/>', 0.70, "0.9238795 0 0.3826834 0"),
  ('<geom type="box" size="0.1 0.15 0.05" mass="2"/>', 0.549, "1 0 0 0"),
  ('<geom type="box" size="0.1 0.15 0.05" mass="2"/>', 0.62, "0.9238795 0.3826834 0 0"),
])
def test_box_primitives_step_parity(geom, z, quat):
  from mjlab_b200.compiler import Spec
  from mjlab_b200.sim import Simulation, SimulationCfg
  from oracle.oracle import Oracle

  m = Spec.from_string(ON_BOX.format(geom=geom, z=z, quat=quat)).compile()
  n = 8
  # capsule-box contact ends are defined to ~1e-3 m (see below): torques, hence accelerations, follow
  pos_tol, acc_tol = (2e-3, 1e-2) if "capsule" in geom else (1e-4, 2e-3)
  sim = Simulation(n, SimulationCfg(), m, "cuda:0")
  sim.set_option("debug_outputs", 1)
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  rng = np.random.default_rng(2)
  qpos = np.tile(m.qpos0, (n, 1))
  qpos[:, 0:2] += rng.uniform(-0.45, 0.45, (n, 2))  # some objects hang over the table edge
  qpos[:, 2] += rng.uniform(-0.01, 0.005, n)
  st = dict(qpos=qpos, qvel=rng.uniform(-0.2, 0.2, (n, 6)))
  load_oracle(o, st)
  load_sim(sim, st)
  for it in range(40):
    o.forward()
    sim.forward()
    torch.cuda.synchronize()
    _check_contacts(sim, o, n, 1e-5)
    e = relerr(T(sim.data.qacc), o.qacc, floor=10.0)
    assert e.max() < acc_tol, (it, e.max())
    k = o.ncon.ravel()
    for w in range(n):
      if k[w]:
        # the ends of a capsule-box contact segment are level-set crossings (d_min + 1e-3 r) of a nearly flat
        # distance profile: their position along the axis is only defined to ~1e-3 m in fp32
        assert np.abs(T(sim.data.contact_pos)[w, : k[w]].ravel() - o.contact_pos[w, : 3 * k[w]]).max() < pos_tol
        assert np.abs(T(sim.data.contact_frame)[w, : k[w]].ravel() - o.contact_frame[w, : 9 * k[w]]).max() < 1e-4
    o.step()
    sim.step()
    # resynchronise so that the comparison stays a one-step comparison
    sim.data.qpos[:] = torch.as_tensor(o.qpos, dtype=torch.float32, device="cuda:0")
    sim.data.qvel[:] = torch.as_tensor(o.qvel, dtype=torch.float32, device="cuda:0")
    sim.data.qacc_warmstart[:] = torch.as_tensor(o.qacc_warmstart, dtype=torch.float32, device="cuda:0")
  assert int(o.ncon.max()) >= 1
  sim.close()


@pytest.mark.parametrize("name,n,spread,dist_tol,acc_tol", [
  ("go1_stairs_small", 128, 1.4, 1e-5, 2e-3),
  ("go1_rough", 192, 2.2, 5e-5, 5e-3),
])
def test_go1_terrain_forward_parity(name, n, spread, dist_tol, acc_tol):
  from mjlab_b200.asset_zoo import load_compiled
  from mjlab_b200.sim import Simulation, SimulationCfg
  from oracle.oracle import Oracle

  m = load_compiled(name)
  sim = Simulation(n, SimulationCfg(), m, "cuda:0")
  sim.set_option("debug_outputs", 1)
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  st = _terrain_states(m, n, 21, spread)
  load_oracle(o, st)
  load_sim(sim, st)
  o.forward()
  sim.forward()
  torch.cuda.synchronize()
  d = sim.data
  same = _check_contacts(sim, o, n, dist_tol)
  assert same == n, same  # candidates are ordered by static index: no dependence on cell borders
  nc = o.ncon.ravel()
  assert nc.max() >= 8 and (nc > 0).mean() > 0.5
  for f in ["xpos", "xmat", "geom_xpos", "geom_xmat", "site_xpos", "subtree_com", "qfrc_bias", "qM"]:
    e = relerr(T(getattr(d, f)).reshape(n, -1), o.field(f).reshape(n, -1)).max()
    assert e < 1e-5, (f, e)
  e = relerr(T(d.qacc), o.qacc, floor=10.0)
  assert e.max() < acc_tol and np.median(e) < acc_tol / 10, (e.max(), np.median(e))
  assert np.abs(T(d.sensordata) - o.sensordata).max() < 1e-3  # foot contact counts
  sim.close()


def test_go1_stairs_rollout_matches_oracle():
  """40 control-free steps from rest on the stairs: state drift stays at fp32 level."""
  from mjlab_b200.asset_zoo import load_compiled
  from mjlab_b200.sim import Simulation, SimulationCfg
  from oracle.oracle import Oracle

  m = load_compiled("go1_stairs_small")
  n = 32
  sim = Simulation(n, SimulationCfg(), m, "cuda:0")
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  st = _terrain_states(m, n, 4, 1.0)
  st["qvel"] *= 0.1
  st["qpos"][:, 2] += 0.03
  load_oracle(o, st)
  load_sim(sim, st)
  for _ in range(40):
    o.step()
  sim.step_n(40)
  torch.cuda.synchronize()
  q, qo = T(sim.data.qpos), o.qpos
  assert np.isfinite(q).all()
  err = np.abs(q - qo).max(axis=1)
  assert np.median(err) < 2e-3 and (err < 2e-2).mean() > 0.85, (np.median(err), err.max())
  sim.close()


def test_go1_rough_full_size_properties():
  """BASELINE config E shape (4096 envs on the 10 x 20 terrain): spawn on the curriculum origins, stand under
  PD control; nothing explodes, the robots stay on their patches, feet report contact."""
  from mjlab_b200.asset_zoo import load_compiled
  from mjlab_b200.sim import Simulation, SimulationCfg
  from mjlab_b200.terrains import env_origins_curriculum

  m = load_compiled("go1_rough")
  n = 4096
  sim = Simulation(n, SimulationCfg(), m, "cuda:0")
  org, level, kind = env_origins_curriculum(n, np.asarray(m.arrays["terrain_origins"]), max_init_level=9)
  key = m.keys["robot/init_state"]
  qpos = np.tile(key["qpos"], (n, 1))
  qpos[:, 0:3] += org
  qpos[:, 2] += 0.02
  sim.data.qpos[:] = torch.as_tensor(qpos, dtype=torch.float32, device="cuda:0")
  sim.data.qvel[:] = 0
  sim.data.ctrl[:] = torch.as_tensor(np.tile(key["ctrl"], (n, 1)), dtype=torch.float32, device="cuda:0")
  sim.step_n(200)
  torch.cuda.synchronize()
  q = T(sim.data.qpos)
  assert np.isfinite(q).all()
  h = q[:, 2] - org[:, 2]
  assert (np.abs(q[:, 0:2] - org[:, 0:2]).max(axis=1) < 0.5).mean() > 0.99
  assert ((h > 0.15) & (h < 0.40)).mean() > 0.99, (h.min(), h.max())
  assert (T(sim.data.sensordata) > 0).mean() > 0.95
  assert int(T(sim.data.overflow).sum()) == 0
  assert np.abs(T(sim.data.qvel)).max() < 1.0
  sim.close()


def test_go1_rough_velocity_env_random_agent():
  """BASELINE config E at env level: Go1 on the rough terrain, random actions, resets on the spawn origins;
  native MDP kernels and the torch implementation agree on the first steps."""
  from mjlab_b200.envs import VelocityEnvCfg, VelocityFlatEnv

  cfg = dict(robot="go1", terrain="rough", num_envs=512, episode_length_s=0.3)
  a = VelocityFlatEnv(VelocityEnvCfg(**cfg), device="cuda:0", native_mdp=False)
  b = VelocityFlatEnv(VelocityEnvCfg(**cfg), device="cuda:0", native_mdp=True)
  assert torch.equal(a.sim.data.qpos[:], b.sim.data.qpos[:])
  z0 = a.sim.data.qpos[:, 2] - a.env_origins[:, 2]
  assert torch.allclose(z0, torch.full_like(z0, float(a.default_qpos[2])), atol=1e-5)
  assert a.env_origins[:, 2].abs().max() > 0.2  # stairs platforms are above / below the flat patches
  g = torch.Generator(device="cuda:0")
  g.manual_seed(5)
  n_trunc = 0
  for k in range(24):
    act = torch.rand((512, 12), generator=g, device="cuda:0") * 2 - 1
    oa, ra, ta, ua, _ = a.step(act)
    ob, rb, tb, ub, _ = b.step(act)
    assert torch.isfinite(oa).all() and torch.isfinite(ra).all()
    ok = (ta == tb) & (ua == ub) & ((ra - rb).abs() < 1e-4) & ((oa - ob).abs().amax(dim=1) < 2e-3)
    assert ok.float().mean() >= 0.97, (k, float(ok.float().mean()))
    for f in ("qpos", "qvel", "qacc_warmstart", "ctrl"):
      getattr(b.sim.data, f)[:] = getattr(a.sim.data, f)[:]
    b.episode_length_buf.copy_(a.episode_length_buf)
    b.last_action.copy_(a.last_action)
    b.command.copy_(a.command)
    n_trunc += int(ua.sum())
  assert n_trunc >= 512  # every env went through a reset onto its terrain origin
  h = a.sim.data.qpos[:, 2] - a.env_origins[:, 2]
  assert (h > 0.05).float().mean() > 0.95 and (h < 0.6).all()  # nobody fell through the terrain
  assert int(a.sim.data.overflow[:].sum()) == 0
  a.close()
  b.close()


def test_obstacle_course_step_parity():
  """Plane + 24 static boxes / spheres / capsules (grid) + 9 free primitives (pair table, all six primitive
  pair types): contact sets identical, accelerations within fp32 tolerance, step by step with resync."""
  from test_boxes_terrain import obstacle_course_xml

  from mjlab_b200.compiler import Spec
  from mjlab_b200.sim import Simulation, SimulationCfg
  from oracle.oracle import Oracle

  m = Spec.from_string(obstacle_course_xml()).compile()
  assert int(m.nstatic) == 24 and int(m.npair) > 0
  n = 16
  sim = Simulation(n, SimulationCfg(nconmax=96 * n), m, "cuda:0")
  sim.set_option("debug_outputs", 1)
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  rng = np.random.default_rng(1)
  q = np.tile(m.qpos0, (n, 1))
  q += rng.uniform(-0.08, 0.08, q.shape) * (np.arange(q.shape[1]) % 7 < 3)
  st = dict(qpos=q, qvel=rng.uniform(-0.3, 0.3, (n, int(m.nv))))
  load_oracle(o, st)
  load_sim(sim, st)
  seen, worst = 0, 0.0
  for it in range(120):
    o.forward()
    sim.forward()
    torch.cuda.synchronize()
    _check_contacts(sim, o, n, 2e-5)
    e = relerr(T(sim.data.qacc), o.qacc, floor=10.0)
    worst = max(worst, float(e.max()))
    seen = max(seen, int(o.ncon.max()))
    o.step()
    sim.step()
    for f in ("qpos", "qvel", "qacc_warmstart"):
      getattr(sim.data, f)[:] = torch.as_tensor(getattr(o, f), dtype=torch.float32, device="cuda:0")
  assert seen >= 8
  assert worst < 1e-2, worst  # capsule-box contact ends, see test_box_primitives_step_parity
  sim.close()
