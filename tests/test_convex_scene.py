"""Mesh and height-field collision inside the fused step kernel (csrc/b2_convex.h): CUDA fp32 against the fp64
oracle on a scene that holds every pair type — on the host emulation of the product sources (CPU suite) and on the
GPU.  The algorithm itself is pinned in tests/test_convex.py."""

import numpy as np
import pytest

from oracle.oracle import Oracle
from util import convex_scene, convex_states, load_oracle, relerr


def _compare(get, o, n, dist_tol, pos_tol):
  nc = o.ncon.ravel()
  assert (get("ncon").ravel() == nc).all(), (get("ncon").ravel(), nc)
  assert (get("nefc").ravel() == o.nefc.ravel()).all()
  cg, og = get("contact_geom").reshape(n, -1, 2), o.contact_geom.reshape(n, -1, 2)
  cp, op = get("contact_pos").reshape(n, -1, 3), o.contact_pos.reshape(n, -1, 3)
  cf, of = get("contact_frame").reshape(n, -1, 9), o.contact_frame.reshape(n, -1, 9)
  types = set()
  for w in range(n):
    k = nc[w]
    assert (cg[w, :k] == og[w, :k]).all()
    assert np.abs(get("contact_dist")[w, :k] - o.contact_dist[w, :k]).max() < dist_tol
    assert np.abs(cp[w, :k] - op[w, :k]).max() < pos_tol, (w, np.abs(cp[w, :k] - op[w, :k]).max())
    assert np.abs(cf[w, :k, :3] - of[w, :k, :3]).max() < 5e-3
    types |= {tuple(p) for p in og[w, :k].tolist()}
  return types


def _pair_types(m, types):
  gt = np.asarray(m.geom_type)
  return {(int(gt[a]), int(gt[b])) for a, b in types}


WANT = {(0, 7), (7, 7), (6, 7), (2, 7), (3, 7), (1, 2), (1, 3), (1, 6), (1, 7)}


def test_emulated_kernel_mesh_and_hfield_contacts():
  from test_kernel_emul import EmulSim, _load

  lib = _load()
  m, anchors, hf = convex_scene()
  n = 24
  sim = EmulSim(lib, m, n, ncon=64)
  o = Oracle(m, nworld=n, maxcon=int(sim.option("maxcon")))
  st = convex_states(m, anchors, hf, n, seed=3)
  load_oracle(o, st)
  sim.load(st)
  o.forward()
  sim.forward()
  types = _compare(sim.field, o, n, 2e-5, 2e-4)
  assert _pair_types(m, types) >= WANT, _pair_types(m, types)
  # (nine free bodies spread over 4 m share one reference point: the small angular accelerations carry the fp32
  # rounding of 2 m lever arms, relative to g that is ~2e-4)
  assert relerr(sim.field("qacc_smooth"), o.qacc_smooth).max() < 5e-4
  assert relerr(sim.field("qacc"), o.qacc).max() < 2e-3
  o.step()
  sim.step(1)
  assert relerr(sim.field("qpos"), o.qpos).max() < 1e-5
  assert relerr(sim.field("qvel"), o.qvel).max() < 2e-3
  sim.close()


@pytest.mark.gpu
def test_gpu_mesh_and_hfield_contacts_match_oracle():
  import torch

  from mjlab_b200.sim import Simulation, SimulationCfg
  from util import load_sim

  m, anchors, hf = convex_scene()
  n = 256
  sim = Simulation(n, SimulationCfg(nconmax=64 * n), m, "cuda:0")
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  st = convex_states(m, anchors, hf, n, seed=11)
  load_oracle(o, st)
  load_sim(sim, st)
  o.forward()
  sim.forward()
  torch.cuda.synchronize()
  get = lambda f: getattr(sim.data, f)[:].cpu().numpy()  # noqa: E731
  # contact sets can differ where a contact sits within fp32 rounding of the margin: compare the worlds that agree
  nc_s, nc_o = get("ncon").ravel(), o.ncon.ravel()
  same = nc_s == nc_o
  assert same.mean() > 0.97, same.mean()
  idx = np.nonzero(same)[0]

  class Sub:
    def __init__(self, o):
      self.o = o

    def __getattr__(self, k):
      return self.o.field(k)[idx]

  types = _compare(lambda f: get(f)[idx], Sub(o), len(idx), 5e-5, 5e-4)
  assert _pair_types(m, types) >= WANT
  e = relerr(get("qacc")[idx], o.qacc[idx])
  assert np.quantile(e, 0.99) < 2e-3 and np.median(e) < 1e-4, (np.median(e), e.max())
  o.step()
  sim.step()
  torch.cuda.synchronize()
  assert relerr(get("qpos")[idx], o.qpos[idx]).max() < 1e-5
  assert np.isfinite(get("qpos")).all()
  # bodies dropped from rest settle: 2 s of simulation, nothing falls through the plane or the height field
  load_sim(sim, convex_states(m, anchors, hf, n, seed=12, drop=True))
  for _ in range(400):
    sim.step()
  torch.cuda.synchronize()
  qp = get("qpos").reshape(n, -1, 7)
  assert np.isfinite(qp).all()
  assert qp[:, :, 2].min() > -0.05, qp[:, :, 2].min()
  lin = get("qvel").reshape(n, -1, 6)[:, :, :3]
  assert np.median(np.abs(lin).max(axis=(1, 2))) < 2.0  # most worlds at rest (spheres may still roll down a slope)
  sim.close()
