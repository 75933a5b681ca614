"""Cylinders and ellipsoids (mjlab_b200/csrc/b2_convex.h: closed-form support points under GJK / EPA, MuJoCo's plane
primitives): the routines against exact values computed another way (the signed distance of two convex bodies from
their support FUNCTIONS, -min_u [h_A(u) + h_B(-u)]; brute force over surface samples for the plane cases), statics
of resting bodies in the oracle,
and kernel-vs-oracle parity of a scene with every pair type on the host emulation (GPU: tests marked gpu)."""

import ctypes

import numpy as np
import pytest

from oracle.oracle import Oracle
from test_convex import G_BOX, G_CAPSULE, G_MESH, G_SPHERE, _call_pair, _lib, _p, _random_shape, _rot
from util import load_oracle, relerr

G_ELLIPSOID, G_CYLINDER = 4, 5


def _surface(t, pos, R, size, n=1500, seed=0):
  """Dense world-frame samples of the surface of a cylinder / ellipsoid."""
  rng = np.random.default_rng(seed)
  if t == G_ELLIPSOID:
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    loc = u * size
  else:
    k = n // 2
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    ring = np.stack([np.cos(ang) * size[0], np.sin(ang) * size[0]], axis=1)
    loc = np.concatenate([np.c_[ring, np.full(k, size[1])], np.c_[ring, np.full(k, -size[1])]])
  return pos + loc @ R.T


def _hfun(t, pos, R, size, verts):
  """Support FUNCTION h(u) = max over the (inflated) shape of x.u, for unit directions u (rows) - closed forms written
  independently of the support POINTS in b2_convex.h."""
  def h(U):
    L = U @ R  # directions in the shape's frame
    base = U @ pos
    if t == G_SPHERE:
      return base + size[0]
    if t == G_CAPSULE:
      return base + size[0] + size[1] * np.abs(L[:, 2])
    if t == G_BOX:
      return base + np.abs(L) @ size
    if t == G_ELLIPSOID:
      return base + np.linalg.norm(L * size, axis=1)
    if t == G_CYLINDER:
      return base + size[0] * np.hypot(L[:, 0], L[:, 1]) + size[1] * np.abs(L[:, 2])
    return base + (L @ verts.T).max(axis=1)
  return h


def _fib(n):
  k = np.arange(n) + 0.5
  phi = np.arccos(1 - 2 * k / n)
  th = np.pi * (1 + 5**0.5) * k
  return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


_DIRS = _fib(20000)


def _signed_distance(hA, hB):
  """-min over unit u of g(u) = h_A(u) + h_B(-u): the separation along the best axis when positive, minus the
  penetration depth (minimum translation) when negative; and the minimiser (points from A to B).  Coarse lattice,
  then a local search on the sphere."""
  from scipy.optimize import minimize

  g = hA(_DIRS) + hB(-_DIRS)
  best_u, best = None, np.inf
  for k in np.argsort(g)[:4]:
    f = lambda x: float(hA((x / np.linalg.norm(x))[None])[0] + hB(-(x / np.linalg.norm(x))[None])[0])  # noqa: E731
    r = minimize(f, _DIRS[k], method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-13, maxiter=3000))
    if r.fun < best:
      best, best_u = r.fun, r.x / np.linalg.norm(r.x)
  return -best, best_u


@pytest.mark.parametrize("prec", [64, 32])
def test_smooth_pairs_match_support_function_optimum(prec):
  """dist (and the normal) of b2c_pair for every pair type with a cylinder or an ellipsoid against the exact
  signed distance of two convex bodies, -min_u [h_A(u) + h_B(-u)]."""
  L, dt, ct = _lib(prec)
  rng = np.random.default_rng(11)
  types = [(G_ELLIPSOID, G_MESH), (G_CYLINDER, G_MESH), (G_SPHERE, G_CYLINDER), (G_CAPSULE, G_ELLIPSOID), (G_CYLINDER, G_BOX),
           (G_ELLIPSOID, G_BOX), (G_ELLIPSOID, G_CYLINDER), (G_CYLINDER, G_CYLINDER), (G_ELLIPSOID, G_ELLIPSOID),
           (G_SPHERE, G_ELLIPSOID), (G_CAPSULE, G_CYLINDER)]
  nsep = npen = nshallow = 0
  worst_sep = worst_pen = 0.0
  for it in range(220):
    t1, t2 = types[it % len(types)]
    s1, v1 = _random_shape(rng, t1)
    s2, v2 = _random_shape(rng, t2)
    R1, R2 = _rot(rng), _rot(rng)
    p1 = rng.normal(size=3) * 0.05
    p2 = p1 + rng.normal(size=3) * rng.choice([0.2, 0.4, 0.6])
    hA, hB = _hfun(t1, p1, R1, s1, v1), _hfun(t2, p2, R2, s2, v2)
    want, u = _signed_distance(hA, hB)
    margin = 0.3
    n, out = _call_pair(L, dt, ct, t1, p1, R1, s1, v1, t2, p2, R2, s2, v2, margin, 1.0)
    if want > margin + 1e-3:
      assert n == 0, (it, want)
      continue
    if want > margin - 1e-3:
      continue
    assert n == 1, (it, t1, t2, want)
    err = float(out[0]) - want
    if want > 0:
      # separated: GJK on exact support points converges to the closest points
      assert abs(err) < (1e-6 if prec == 64 else 2e-4), (it, t1, t2, out[0], want)
      worst_sep = max(worst_sep, abs(err))
      nsep += 1
    else:
      # overlapping: EPA's polytope (<= 28 support points) is inscribed in the smooth Minkowski difference: the
      # depth comes out smaller, by at most a few per cent of the shapes' size
      assert -1e-6 - (2e-4 if prec == 32 else 0) <= err <= 0.012 + 0.03 * abs(want), (it, t1, t2, out[0], want)
      worst_pen = max(worst_pen, err)
      npen += 1
    assert abs(np.linalg.norm(out[4:7]) - 1) < 1e-4
    nn = out[4:7].astype(float)[None]
    if want > 5e-3:
      assert np.dot(nn[0], u) > 0.999, (it, t1, t2, nn, u)  # the separating direction is unique
    elif -0.1 < want < -5e-3:
      # overlapping bodies can have several nearly equal escape directions: the reported one must be (nearly) as
      # short as the best.  (Bodies sunk into each other by more than 0.1 - a third of their size - are left out:
      # 24 expansions do not converge the direction there, only the depth; contacts in a simulation are shallow.)
      g_n = float(hA(nn)[0] + hB(-nn)[0])
      assert g_n <= -want + (0.015 if prec == 64 else 0.02), (it, t1, t2, g_n, want)
      nshallow += 1
  assert nsep > 40 and npen > 40 and nshallow > 12, (nsep, npen, nshallow)
  print(f"worst separated error {worst_sep:.2e}, worst depth deficit {worst_pen:.2e}")


def _plane_call(L, dt, ct, t2, pos, R, size, margin):
  out = np.zeros(28, dtype=dt)
  a = [np.zeros(3, dtype=dt), np.array([0, 0, 1], dtype=dt), np.ascontiguousarray(pos, dtype=dt),
       np.ascontiguousarray(R.reshape(-1), dtype=dt), np.ascontiguousarray(size, dtype=dt)]
  n = L.b2o_prim_plane_smooth(_p(a[0]), _p(a[1]), ctypes.c_int(t2), _p(a[2]), _p(a[3]), _p(a[4]), ct(margin), _p(out))
  return n, out.reshape(4, 7)


@pytest.mark.parametrize("prec", [64, 32])
def test_plane_cylinder_and_ellipsoid(prec):
  L, dt, ct = _lib(prec)
  rng = np.random.default_rng(3)
  tol = 1e-9 if prec == 64 else 2e-6
  for it in range(60):
    R = _rot(rng)
    size = rng.uniform(0.05, 0.3, size=3)
    pos = np.array([rng.normal() * 0.3, rng.normal() * 0.3, rng.uniform(0.0, 0.35)])
    for t in (G_ELLIPSOID, G_CYLINDER):
      low = _surface(t, pos, R, size, n=20000, seed=it)[:, 2].min()  # brute-force lowest point (samples: from above)
      n, c = _plane_call(L, dt, ct, t, pos, R, size, 0.02)
      if low > 0.02 + 1e-3:
        assert n == 0
        continue
      if low > 0.02 - 1e-3:
        continue
      assert n >= 1
      assert -1e-3 * size.max() - 2e-4 <= low - c[0, 0] + 0 and low - c[0, 0] <= 2e-3, (it, t, low, c[0, 0])  # deepest point
      assert np.allclose(c[:n, 4:7], [0, 0, 1], atol=1e-6)
      assert np.allclose(c[:n, 3], c[:n, 0] * 0.5, atol=1e-6)  # position midway between the surface point and the plane
      if t == G_ELLIPSOID:
        assert n == 1
  I = np.eye(3)
  # cylinder standing on a cap, 1 mm into the plane: a triangle of rim points, all at the same depth
  n, c = _plane_call(L, dt, ct, G_CYLINDER, np.array([0.0, 0.0, 0.199]), I, np.array([0.1, 0.2, 0.0]), 0.0)
  assert n == 3 and np.allclose(c[:3, 0], -0.001, atol=tol)
  xy = c[:3, 1:3]
  assert np.allclose(np.linalg.norm(xy, axis=1), 0.1, atol=1e-6) and np.allclose(xy.sum(0), 0, atol=1e-6)
  # lying on its side: both ends of the lowest generator, the near-cap triangle points stay above the margin
  Rx = np.array([[1.0, 0, 0], [0, 0, -1], [0, 1, 0]])  # local z -> world -y
  n, c = _plane_call(L, dt, ct, G_CYLINDER, np.array([0.0, 0.0, 0.099]), Rx, np.array([0.1, 0.2, 0.0]), 0.0)
  assert n == 2 and np.allclose(c[:2, 0], -0.001, atol=tol) and np.allclose(sorted(c[:2, 2]), [-0.2, 0.2], atol=1e-6)
  # tilted 45 degrees, rim touching: one contact
  s = np.sqrt(0.5)
  Rt = np.array([[1.0, 0, 0], [0, s, -s], [0, s, s]])
  zc = s * (0.1 + 0.2)
  n, c = _plane_call(L, dt, ct, G_CYLINDER, np.array([0.0, 0.0, zc - 0.002]), Rt, np.array([0.1, 0.2, 0.0]), 0.0)
  assert n == 1 and abs(c[0, 0] + 0.002) < 1e-6
  # ellipsoid resting on its short axis
  n, c = _plane_call(L, dt, ct, G_ELLIPSOID, np.array([0.3, 0.1, 0.049]), I, np.array([0.2, 0.1, 0.05]), 0.0)
  assert n == 1 and abs(c[0, 0] + 0.001) < 1e-6 and np.allclose(c[0, 1:3], [0.3, 0.1], atol=1e-6)


def smooth_scene():
  """Ground plane, a wavy height field, a static box, and free bodies with cylinder / ellipsoid geoms next to the
  other shapes: plane-cyl/ell, sphere-cyl, capsule-ell, cyl-box, ell-box, ell-cyl, cyl-mesh, hfield-cyl/ell."""
  from mjlab_b200.compiler.spec import Spec

  rng = np.random.default_rng(17)
  spec = Spec()
  spec.option.timestep = 0.005
  spec.add_mesh("poly", vertex=rng.normal(size=(12, 3)) * np.array([0.12, 0.1, 0.08]))
  nrow, ncol = 12, 14
  ii, jj = np.meshgrid(np.arange(nrow), np.arange(ncol), indexing="ij")
  elev = np.clip(0.5 + 0.3 * np.sin(0.7 * ii) * np.cos(0.5 * jj), 0, 1)
  spec.add_hfield("waves", size=[1.4, 1.2, 0.2, 0.1], nrow=nrow, ncol=ncol, userdata=elev)
  wb = spec.worldbody
  wb.add_geom(name="floor", type="plane", size=[0, 0, 0.05])
  wb.add_geom(name="terrain", type="hfield", hfieldname="waves", pos=[4.0, 0.0, 0.0])
  anchors = {}

  def body(name, pos, **geom):
    b = wb.add_body(name=name, pos=pos)
    b.add_freejoint(name=name + "_j")
    b.add_geom(name=name + "_g", density=700.0, **geom)
    anchors[name] = np.array(pos, dtype=float)

  body("cy1", [0.0, 0.0, 0.16], type="cylinder", size=[0.08, 0.12])
  body("el1", [0.22, 0.0, 0.14], type="ellipsoid", size=[0.12, 0.08, 0.06])
  body("sp", [0.0, 0.2, 0.15], type="sphere", size=[0.07])
  body("cp", [0.25, 0.2, 0.16], type="capsule", size=[0.04, 0.1])
  body("bx", [-0.22, 0.0, 0.15], type="box", size=[0.08, 0.06, 0.1])
  body("ms", [0.0, -0.22, 0.16], type="mesh", meshname="poly")
  body("cy2", [0.24, -0.2, 0.15], type="cylinder", size=[0.06, 0.05])
  body("hcy", [3.7, -0.2, 0.0], type="cylinder", size=[0.07, 0.1])
  body("hel", [4.3, 0.3, 0.0], type="ellipsoid", size=[0.1, 0.07, 0.05])
  m = spec.compile()
  hf = dict(pos=np.array([4.0, 0.0, 0.0]), size=np.array([1.4, 1.2, 0.2, 0.1]), data=elev)
  return m, anchors, hf


def test_compiler_accepts_smooth_shapes_and_their_inertia():
  m, _, _ = smooth_scene()
  gt = np.asarray(m.geom_type)
  assert (gt == G_CYLINDER).sum() == 3 and (gt == G_ELLIPSOID).sum() == 2
  names = m.names["body"]
  bm = np.asarray(m.body_mass)
  assert bm[names.index("cy1")] == pytest.approx(700.0 * np.pi * 0.08**2 * 0.24, rel=1e-6)
  assert bm[names.index("el1")] == pytest.approx(700.0 * 4 / 3 * np.pi * 0.12 * 0.08 * 0.06, rel=1e-6)
  pairs = {(int(gt[a]), int(gt[b])) for a, b in zip(np.asarray(m.pair_geom1), np.asarray(m.pair_geom2))}
  assert {(0, 4), (0, 5), (1, 4), (1, 5), (2, 5), (3, 4), (4, 5), (5, 6), (4, 6), (5, 7), (5, 5)} <= pairs


def _states(m, anchors, hf, n, seed):
  from util import convex_states

  return convex_states(m, anchors, hf, n, seed)


def test_oracle_statics_cylinder_and_ellipsoid_rest_on_the_plane():
  """A cylinder lying on its side and an ellipsoid on its short axis come to rest on the plane: the contact forces
  carry the weight, the bodies stop."""
  from mjlab_b200.compiler.spec import Spec

  spec = Spec()
  spec.option.timestep = 0.002
  wb = spec.worldbody
  wb.add_geom(name="floor", type="plane", size=[0, 0, 0.05])
  b = wb.add_body(name="cyl", pos=[0, 0, 0.1], quat=[np.sqrt(0.5), np.sqrt(0.5), 0, 0])
  b.add_freejoint(name="cj")
  b.add_geom(name="cg", type="cylinder", size=[0.1, 0.2], density=500.0)
  e = wb.add_body(name="ell", pos=[1.0, 0, 0.05])
  e.add_freejoint(name="ej")
  e.add_geom(name="eg", type="ellipsoid", size=[0.2, 0.15, 0.05], density=500.0)
  m = spec.compile()
  o = Oracle(m, nworld=1)
  for _ in range(600):
    o.step()
  assert np.abs(o.qvel).max() < 2e-3
  assert o.qpos[0, 2] == pytest.approx(0.1, abs=2e-3) and o.qpos[0, 9] == pytest.approx(0.05, abs=2e-3)
  o.forward()
  w = float(np.asarray(m.body_mass).sum()) * 9.81
  f = o.contact_force.reshape(-1, 3)[: int(o.ncon[0]), 0].sum()
  assert f == pytest.approx(w, rel=2e-3)
  assert int(o.ncon[0]) == 3  # two generator ends of the cylinder + the ellipsoid's support point


def test_emulated_kernel_smooth_scene_parity():
  """Every cylinder / ellipsoid pair type through the fp32 kernel (host emulation) against the fp64 oracle."""
  from test_kernel_emul import EmulSim, _load

  lib = _load()
  m, anchors, hf = smooth_scene()
  n = 24
  sim = EmulSim(lib, m, n, ncon=48)
  o = Oracle(m, nworld=n, maxcon=int(sim.option("maxcon")))
  st = _states(m, anchors, hf, n, 4)
  load_oracle(o, st)
  sim.load(st)
  o.forward()
  sim.forward()
  nc = o.ncon.ravel()
  gt = np.asarray(m.geom_type)
  og = o.contact_geom.reshape(n, -1, 2)
  seen = set()
  same = sim.field("ncon").ravel() == nc
  assert same.mean() > 0.9, same
  for w in np.nonzero(same)[0]:
    k = nc[w]
    assert (sim.field("contact_geom")[w, :k] == og[w, :k]).all()
    if k:
      # smooth shapes: EPA works on a polytope of <= 28 support points and fp32 stops earlier than fp64 (a few mm on
      # overlaps of several cm in these random states)
      assert np.abs(sim.field("contact_dist")[w, :k] - o.contact_dist[w, :k]).max() < 5e-3
      seen |= {(int(gt[a]), int(gt[b])) for a, b in og[w, :k]}
  assert {(0, 4), (0, 5), (1, 4), (1, 5)} <= seen and len(seen) >= 8, seen
  # (accelerations inherit the mm differences of the deep EPA contacts through the contact stiffness)
  assert np.median(relerr(sim.field("qacc")[same], o.qacc[same])) < 1e-2
  sim.close()


@pytest.mark.gpu
def test_gpu_smooth_scene_parity():
  """The smooth-shape scene on the GPU against the fp64 oracle.  Contact sets must agree; depths are compared by
  quantiles: on overlaps of several centimetres (these random states) EPA in fp32 occasionally stops early on a curved
  body - 2.3 cm on one of ~2500 contacts in the r02 run, where the host emulation of the same code converged - so
  the bound on the worst contact is the bodies' size, and the bounds that matter are the median and the 90th
  percentile."""
  import torch

  from mjlab_b200.sim import Simulation, SimulationCfg
  from util import convex_states, load_sim

  m, anchors, hf = smooth_scene()
  n = 256
  sim = Simulation(n, SimulationCfg(nconmax=48 * n), m, "cuda:0")
  o = Oracle(m, nworld=n, maxcon=int(sim.get_option("maxcon")))
  st = _states(m, anchors, hf, n, 6)
  load_oracle(o, st)
  load_sim(sim, st)
  o.forward()
  sim.forward()
  torch.cuda.synchronize()
  get = lambda f: getattr(sim.data, f)[:].cpu().numpy()  # noqa: E731
  nc = o.ncon.ravel()
  same = get("ncon").ravel() == nc
  assert same.mean() > 0.9, same.mean()
  errs = []
  for w in np.nonzero(same)[0]:
    k = nc[w]
    assert (get("contact_geom")[w].reshape(-1, 2)[:k] == o.contact_geom[w].reshape(-1, 2)[:k]).all()
    errs += list(np.abs(get("contact_dist")[w, :k] - o.contact_dist[w, :k]))
  errs = np.array(errs)
  assert len(errs) > 1000
  assert np.median(errs) < 1e-5 and np.quantile(errs, 0.9) < 5e-3 and errs.max() < 0.1, (np.median(errs), np.quantile(errs, 0.9), errs.max())
  assert np.median(relerr(get("qacc")[same], o.qacc[same])) < 2e-2
  # dropped from rest: everything lands and stays on the ground
  load_sim(sim, convex_states(m, anchors, hf, n, 9, drop=True))
  sim.forward()
  for _ in range(200):
    sim.step()
  torch.cuda.synchronize()
  assert torch.isfinite(sim.data.qpos[:]).all() and torch.isfinite(sim.data.qvel[:]).all()
  assert float(sim.data.xpos[:, 1:, 2].min()) > -0.05
  sim.close()


def test_emulated_kernel_smooth_scene_drop_stays_bounded():
  """Bodies dropped from rest onto the plane and the height field through the fp32 kernel (host emulation): the
  cylinders and ellipsoids land, roll and slow down - nothing tunnels through the ground or gains energy."""
  from test_kernel_emul import EmulSim, _load
  from util import convex_states

  lib = _load()
  m, anchors, hf = smooth_scene()
  n = 6
  sim = EmulSim(lib, m, n, ncon=48)
  st = convex_states(m, anchors, hf, n, 9, drop=True)
  sim.load(st)
  sim.forward()
  z0 = sim.field("xpos")[:, 1:, 2].copy()
  vmax = 0.0
  for _ in range(8):
    sim.step(25)
    assert np.isfinite(sim.field("qpos")).all()
    vmax = max(vmax, float(np.abs(sim.field("qvel")).max()))
  z = sim.field("xpos")[:, 1:, 2]
  assert (z > -0.02).all() and (z <= z0 + 1e-3).all()  # on or above the ground, never above the drop height
  assert vmax < 25.0 and np.abs(sim.field("qvel")[:, :3]).max() < 6.0
  assert int(sim.field("ncon").min()) >= 4  # everything has landed
  sim.close()


@pytest.mark.parametrize("prec", [64, 32])
def test_shallow_overlaps_are_accurate(prec):
  """The regime a simulation lives in: bodies overlapping by 0.2 mm - 3 cm.  There the expanding polytope converges:
  depth within 1e-5 m of the exact value in fp32 (2e-6 in fp64)."""
  L, dt, ct = _lib(prec)
  rng = np.random.default_rng(21)
  types = [(G_ELLIPSOID, G_MESH), (G_CYLINDER, G_MESH), (G_SPHERE, G_CYLINDER), (G_CAPSULE, G_ELLIPSOID), (G_CYLINDER, G_BOX),
           (G_ELLIPSOID, G_BOX), (G_ELLIPSOID, G_CYLINDER), (G_CYLINDER, G_CYLINDER), (G_ELLIPSOID, G_ELLIPSOID)]
  coarse = _fib(2000)
  errs = []
  for it in range(400):
    t1, t2 = types[it % len(types)]
    s1, v1 = _random_shape(rng, t1)
    s2, v2 = _random_shape(rng, t2)
    R1, R2 = _rot(rng), _rot(rng)
    p1 = np.zeros(3)
    along = rng.normal(size=3)
    along /= np.linalg.norm(along)
    hA = _hfun(t1, p1, R1, s1, v1)
    target = -rng.uniform(0.0005, 0.02)
    lo, hi = 0.0, 2.0  # bisection on the centre distance for an overlap of about `target`
    for _ in range(30):
      mid = 0.5 * (lo + hi)
      hB = _hfun(t2, p1 + along * mid, R2, s2, v2)
      if -(hA(coarse) + hB(-coarse)).min() < target:
        lo = mid
      else:
        hi = mid
    p2 = p1 + along * 0.5 * (lo + hi)
    want, _ = _signed_distance(hA, _hfun(t2, p2, R2, s2, v2))
    if not -0.03 < want < -2e-4:
      continue
    n, out = _call_pair(L, dt, ct, t1, p1, R1, s1, v1, t2, p2, R2, s2, v2, 0.0, 1.0)
    assert n == 1, (it, t1, t2, want)
    errs.append(abs(float(out[0]) - want))
    if len(errs) >= 100:
      break
  errs = np.array(errs)
  assert len(errs) >= 80
  assert errs.max() < (2e-6 if prec == 64 else 1e-5), errs.max()
