"""Height-field sub-terrains (mjlab_b200/terrains.py restating terrains/heightfield_terrains.py) and Go1 walking
surfaces made of them: generator invariants, oracle statics, kernel-vs-oracle parity on the host emulation and on
the GPU."""

import numpy as np
import pytest

from mjlab_b200 import terrains as T
from oracle.oracle import Oracle
from util import hfield_states, load_oracle, relerr, surface_height


def test_hf_generators_shapes_and_levels():
  size = (8.0, 8.0)
  hf, org = T.hf_pyramid_sloped(size, 0.5, slope_range=(0.0, 1.0), platform_width=2.0, border_width=0.25)
  assert hf.data.shape == (80, 80) and hf.data.min() == 0 and hf.data.max() == 1
  assert np.allclose(hf.size[:2], [4, 4]) and hf.size[2] > 0 and hf.size[3] == pytest.approx(hf.size[2])
  # border rows stay at level 0, the platform is the plateau at the top
  assert (hf.data[:2] == 0).all() and (hf.data[:, -2:] == 0).all()
  assert (hf.data[35:45, 35:45] == 1).all()
  assert org[2] == pytest.approx(hf.size[2] - hf.size[3]) and np.allclose(hf.pos, [4, 4, 0])
  # inner region 76 samples: peak level int(0.5 * 7.6 / 2 / 0.005) = 380, product profile clipped at the platform
  # corner (28/38 of the way up on both axes): 380 * (28/38)^2 = 206 levels of 5 mm
  assert hf.size[2] == pytest.approx(206 * 0.005)
  inv, org_i = T.hf_pyramid_sloped(size, 0.5, slope_range=(0.0, 1.0), platform_width=2.0, border_width=0.25, inverted=True)
  assert (inv.data[35:45, 35:45] == 0).all() and inv.pos[2] == pytest.approx(-inv.size[2])
  assert np.allclose(inv.data, 1 - hf.data)
  rng = np.random.default_rng(0)
  rr, org_r = T.hf_random_uniform(size, rng, noise_range=(0.02, 0.10), noise_step=0.02, border_width=0.25)
  assert rr.data.shape == (80, 80) and org_r[2] == pytest.approx(0.06)
  lv = np.unique(np.round(rr.data * rr.size[2] / 0.005).astype(int))
  assert len(lv) > 5  # spline-interpolated levels in vertical_scale units
  wv, org_w = T.hf_wave(size, 1.0, amplitude_range=(0.0, 0.2), num_waves=4, border_width=0.25)
  assert wv.size[2] == pytest.approx(2 * 2 * int(0.5 * 0.2 / 0.005) * 0.005, abs=0.011)
  assert wv.pos[2] == pytest.approx(-wv.size[2] / 2) and org_w[2] == 0
  assert wv.size[3] == pytest.approx(0.25 * wv.size[2])
  with pytest.raises(ValueError):
    T.hf_wave(size, 1.0, amplitude_range=(0, 0.2), num_waves=0)
  with pytest.raises(ValueError):
    T.hf_wave(size, 1.0, amplitude_range=(0, 0.2), border_width=0.05)


def test_full_terrain_mixes_boxes_and_height_fields():
  cfg = T.RoughTerrainCfg(num_rows=3, num_cols=14, border_width=1.0, sub_terrains=T.FULL_SUB_TERRAINS)
  items, origins = T.generate_terrain(cfg)
  nhf = sum(isinstance(i, T.HeightField) for i in items)
  assert origins.shape == (3, 14, 3)
  # proportions .4 .3 .3 | .1 .1 .2 .2 (sum 1.6): columns with c/14 >= 0.625 are height fields -> 5 columns x 3 rows
  assert nhf == 15
  sp, _ = T.terrain_spec(cfg)
  assert len(sp.hfields) == 15


def test_oracle_go1_stands_on_height_fields():
  from mjlab_b200.asset_zoo import load_compiled

  m = load_compiled("go1_hf_small")
  assert int(m.npair) > 0 and (np.asarray(m.geom_type) == 1).sum() == 8
  n = 8
  o = Oracle(m, nworld=n, maxcon=48)
  st = hfield_states(m, n, 3, 1.2, clearance=(0.1, 0.15))
  load_oracle(o, st)
  o.ctrl[:] = st["qpos"][:, 7:]
  for _ in range(300):
    o.step()
  assert np.isfinite(o.qpos).all()
  gt = np.asarray(m.geom_type)
  touching = 0
  for w in range(n):
    g = o.contact_geom[w].reshape(-1, 2)[: o.ncon[w, 0]]
    touching += int((gt[g[:, 0]] == 1).any())
  assert touching >= n - 1  # robots rest on height-field prisms (one may have walked onto the border boxes)
  # nobody fell through: the trunk stays above the surface under it
  for w in range(n):
    assert o.qpos[w, 2] > surface_height(m, o.qpos[w, 0], o.qpos[w, 1]) - 0.3 + 0.1, w
  # most robots are at rest (one on a steep pyramid face may still be sliding)
  assert np.median(np.abs(o.qvel).max(axis=1)) < 0.3 and np.abs(o.qvel).max() < 3.0


def test_emulated_kernel_on_height_field_terrain():
  from mjlab_b200.asset_zoo import load_compiled
  from test_kernel_emul import EmulSim, _load

  lib = _load()
  m = load_compiled("go1_hf_small")
  n = 32
  sim = EmulSim(lib, m, n, ncon=48)
  o = Oracle(m, nworld=n, maxcon=int(sim.option("maxcon")))
  st = hfield_states(m, n, 21, 1.5)
  load_oracle(o, st)
  sim.load(st)
  o.forward()
  sim.forward()
  nc = o.ncon.ravel()
  assert nc.sum() >= 8
  assert (sim.field("ncon").ravel() == nc).all()
  cg, og = sim.field("contact_geom"), o.contact_geom.reshape(n, -1, 2)
  for w in range(n):
    assert (cg[w, : nc[w]] == og[w, : nc[w]]).all()
    if nc[w]:
      assert np.abs(sim.field("contact_dist")[w, : nc[w]] - o.contact_dist[w, : nc[w]]).max() < 2e-5
  assert relerr(sim.field("qacc"), o.qacc).max() < 2e-3
  # the fields are welded to the world: the engine poses them once at create (no per-step work, no shared-memory
  # slot) - every world's geom_xpos / geom_xmat rows must still hold them
  hf = np.nonzero(np.asarray(m.geom_type) == 1)[0]
  assert np.abs(sim.field("geom_xpos")[:, hf] - o.geom_xpos.reshape(n, -1, 3)[:, hf]).max() < 1e-6
  assert np.abs(sim.field("geom_xmat").reshape(n, -1, 9)[:, hf] - o.geom_xmat.reshape(n, -1, 9)[:, hf]).max() < 1e-6
  assert sim.option("smem_bytes_per_env") < 13.5 * 1024  # (no pose slots for the fields; on the 70-field bench terrain the slots were 3.9 KB of 15.6 KB: 12 instead of 16 warps per SM)
  o.step()
  sim.step(1)
  assert relerr(sim.field("qpos"), o.qpos).max() < 1e-5
  sim.close()


def test_emulated_kernel_robots_straddling_field_borders():
  """Robots on the seams and corners of the terrain's grid of height fields: the hierarchical broadphase
  (fields under the robot first, then their pairs) finds candidates on two or four fields and must hand them to
  the narrowphase in pair-table order; the lane-distributed narrowphase then merges the prisms of several pairs
  per round.  Contact sets, order and depths against the oracle."""
  from mjlab_b200.asset_zoo import load_compiled
  from test_kernel_emul import EmulSim, _load
  from util import make_states, surface_height

  lib = _load()
  m = load_compiled("go1_rough_hf")
  gt = np.asarray(m.geom_type)
  hf = np.nonzero(gt == 1)[0]
  ctr = np.asarray(m.geom_pos)[hf] + np.asarray(m.body_pos)[np.asarray(m.geom_bodyid)[hf]]
  half = np.asarray(m.hfield_size)[np.asarray(m.geom_dataid)[hf], 0]
  rng = np.random.default_rng(4)
  n = 24
  st = make_states(m, n, seed=12, z_range=(0.0, 0.0))
  key = make_states(m, 1, seed=0, z_range=(0.0, 0.0))["qpos"][0, 2]
  for w in range(n):
    k = rng.integers(0, len(hf))
    x = ctr[k, 0] + half[k] * (1 if w % 2 else -1) + rng.uniform(-0.08, 0.08)  # on the x seam
    y = ctr[k, 1] + (half[k] * (1 if w % 4 < 2 else -1) + rng.uniform(-0.08, 0.08) if w % 3 else rng.uniform(-2, 2))
    st["qpos"][w, 0:2] = (x, y)
    st["qpos"][w, 2] = max(surface_height(m, x + dx, y + dy) for dx in (-0.2, 0.2) for dy in (-0.2, 0.2)) + key - 0.04
  sim = EmulSim(lib, m, n, ncon=48)
  o = Oracle(m, nworld=n, maxcon=int(sim.option("maxcon")))
  load_oracle(o, st)
  sim.load(st)
  o.forward()
  sim.forward()
  nc = o.ncon.ravel()
  og = o.contact_geom.reshape(n, -1, 2)
  fields_touched = [len({int(g) for g in og[w, : nc[w], 0] if gt[g] == 1}) for w in range(n)]
  assert sum(f >= 2 for f in fields_touched) >= 4, fields_touched  # the seams are exercised
  assert (sim.field("ncon").ravel() == nc).all()
  cg = sim.field("contact_geom")
  for w in range(n):
    assert (cg[w, : nc[w]] == og[w, : nc[w]]).all(), w
    if nc[w]:
      assert np.abs(sim.field("contact_dist")[w, : nc[w]] - o.contact_dist[w, : nc[w]]).max() < 2e-5
  sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("go1_hf_small", 256), ("go1_rough_hf", 192)])
def test_gpu_height_field_terrain_parity(name, n):
  import torch

  from mjlab_b200.asset_zoo import load_compiled
  from mjlab_b200.sim import Simulation, SimulationCfg
  from util import load_sim

  m = load_compiled(name)
  sim = Simulation(n, SimulationCfg(nconmax=48 * n), m, "cuda:0")
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  st = hfield_states(m, n, 9, 1.5)
  load_oracle(o, st)
  load_sim(sim, st)
  o.forward()
  sim.forward()
  torch.cuda.synchronize()
  get = lambda f: getattr(sim.data, f)[:].cpu().numpy()  # noqa: E731
  nc_s, nc_o = get("ncon").ravel(), o.ncon.ravel()
  same = nc_s == nc_o
  assert same.mean() > 0.97, same.mean()
  gt = np.asarray(m.geom_type)
  hf_contacts = 0
  for w in np.nonzero(same)[0]:
    k = nc_o[w]
    a, b = get("contact_geom")[w].reshape(-1, 2)[:k], o.contact_geom[w].reshape(-1, 2)[:k]
    assert (a == b).all()
    if k:
      assert np.abs(get("contact_dist")[w, :k] - o.contact_dist[w, :k]).max() < 1e-4
      hf_contacts += int((gt[b[:, 0]] == 1).sum())
  assert hf_contacts > n // 4
  e = relerr(get("qacc")[same], o.qacc[same])
  assert np.median(e) < 1e-4 and np.quantile(e, 0.99) < 5e-3, (np.median(e), e.max())
  for _ in range(200):
    sim.step()
  torch.cuda.synchronize()
  assert np.isfinite(get("qpos")).all()
  sim.close()
