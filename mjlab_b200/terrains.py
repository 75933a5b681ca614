"""Procedural terrains: box sub-terrains (BASELINE config E: Go1 rough terrain) and height-field sub-terrains.

Geometry-equivalent restatement of the reference's box sub-terrains and grid generator —
``src/mjlab/terrains/primitive_terrains.py:53-376`` (flat patch, pyramid stairs, inverted pyramid stairs),
``terrains/utils.py:11-105`` (finite plane, four-box border) and ``terrains/terrain_generator.py:62-248`` (grid
placement, curriculum difficulty per row, sub-terrain type per column, outer border) — producing the same
boxes in the same order (ids leak into contact/geom indices), written as data-in/data-out numpy code:
every function returns ``[(half_size, centre), ...]`` instead of mutating an MjSpec.

The physics engine treats these geoms as *static* (world-welded) and finds candidates through a uniform grid
built at compile time (``compiler/compile.py``), not through the pair table.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from mjlab_b200.compiler import spec as S

Box = tuple  # (half_size(3), centre(3))


def ring_boxes(outer, inner, height, centre) -> list[Box]:
  """Hollow rectangle as four boxes: +y strip, -y strip (full width), +x, -x (inner height)
  — the order of ``make_border`` (``terrains/utils.py:36-105``)."""
  tx, ty = 0.5 * (outer[0] - inner[0]), 0.5 * (outer[1] - inner[1])
  cx, cy, cz = centre
  return [
    (np.array([outer[0], ty, height]) / 2, np.array([cx, cy + inner[1] / 2 + ty / 2, cz])),
    (np.array([outer[0], ty, height]) / 2, np.array([cx, cy - inner[1] / 2 - ty / 2, cz])),
    (np.array([tx, inner[1], height]) / 2, np.array([cx + inner[0] / 2 + tx / 2, cy, cz])),
    (np.array([tx, inner[1], height]) / 2, np.array([cx - inner[0] / 2 - tx / 2, cy, cz])),
  ]


def flat_patch(size, thickness: float = 1.0):
  """One box whose top face is z = 0 (``make_plane(center_zero=False)``, ``utils.py:11-33``)."""
  box = (np.array([size[0], size[1], thickness]) / 2, np.array([size[0] / 2, size[1] / 2, -thickness / 2]))
  return [box], np.array([size[0] / 2, size[1] / 2, 0.0])


def pyramid_stairs(size, difficulty, step_height_range=(0.0, 0.1), step_width=0.3, platform_width=3.0,
                   border_width=1.0, inverted=False):
  """Concentric square steps going up (or, inverted, down) towards a central platform
  (``primitive_terrains.py:66-224`` / ``:227-376``).  Returns (boxes, spawn origin)."""
  h = step_height_range[0] + difficulty * (step_height_range[1] - step_height_range[0])
  nx = (size[0] - 2 * border_width - platform_width) // (2 * step_width) + 1
  ny = (size[1] - 2 * border_width - platform_width) // (2 * step_width) + 1
  ns = int(min(nx, ny))
  total = (ns + 1) * h
  cx, cy = size[0] / 2, size[1] / 2
  inner = (size[0] - 2 * border_width, size[1] - 2 * border_width)
  boxes: list[Box] = []
  if border_width > 0:
    boxes += ring_boxes(size, inner, h, (cx, cy, -h / 2))
  for k in range(ns):
    bs = (inner[0] - 2 * k * step_width, inner[1] - 2 * k * step_width)
    off = (k + 0.5) * step_width
    if inverted:
      z, bh = -total / 2 - (k + 1) * h / 2, total - (k + 1) * h
    else:
      z, bh = k * h / 2, (k + 2) * h
    long_x = np.array([bs[0], step_width, bh]) / 2
    boxes.append((long_x, np.array([cx, cy + inner[1] / 2 - off, z])))
    boxes.append((long_x, np.array([cx, cy - inner[1] / 2 + off, z])))
    long_y = np.array([step_width, bs[1] - 2 * step_width, bh]) / 2
    boxes.append((long_y, np.array([cx + inner[0] / 2 - off, cy, z])))
    boxes.append((long_y, np.array([cx - inner[0] / 2 + off, cy, z])))
  mid = (inner[0] - 2 * ns * step_width, inner[1] - 2 * ns * step_width)
  if inverted:
    boxes.append((np.array([mid[0], mid[1], h]) / 2, np.array([cx, cy, -total - h / 2])))
    origin = np.array([cx, cy, -total])
  else:
    boxes.append((np.array([mid[0], mid[1], (ns + 2) * h]) / 2, np.array([cx, cy, ns * h / 2])))
    origin = np.array([cx, cy, (ns + 1) * h])
  return boxes, origin


@dataclass
class HeightField:
  """One ``<hfield>`` sub-terrain: ``data[nrow, ncol]`` in [0, 1], ``size = (rx, ry, elevation, base)`` and the
  geom position (patch frame; the generator adds the patch corner)."""

  data: np.ndarray
  size: np.ndarray
  pos: np.ndarray


def _hf_finish(levels: np.ndarray, size, vertical_scale, base_ratio, zpos):
  """Integer height levels -> normalised samples + hfield size (``heightfield_terrains.py:196-225``: elevation
  = (max - min) levels x vertical_scale, base thickness = elevation x ratio; a constant field gets range 1)."""
  lo, hi = int(levels.min()), int(levels.max())
  rng_lv = hi - lo if hi != lo else 1
  elevation = rng_lv * vertical_scale
  data = (levels - lo) / rng_lv
  hsize = np.array([size[0] / 2, size[1] / 2, elevation, elevation * base_ratio])
  return data, hsize, elevation


def _hf_inner(size, horizontal_scale, border_width):
  """Sample counts of the patch and of its inner region (a flat border of ``border_width`` stays at level 0)."""
  if 0 < border_width < horizontal_scale:
    raise ValueError(f"Border width ({border_width}) must be >= horizontal scale ({horizontal_scale})")
  bp = int(border_width / horizontal_scale)
  npx, npy = int(size[0] / horizontal_scale), int(size[1] / horizontal_scale)
  return bp, npx, npy, npx - 2 * bp, npy - 2 * bp


def hf_pyramid_sloped(size, difficulty, slope_range, platform_width=1.0, inverted=False, border_width=0.0,
                      horizontal_scale=0.1, vertical_scale=0.005, base_thickness_ratio=1.0):
  """Pyramid with a flat top (or, inverted, a pit): height = h_max * tx * ty with tx, ty the triangle profiles
  across the inner region, clipped at the platform's height (``heightfield_terrains.py:103-252``)."""
  slope = slope_range[0] + difficulty * (slope_range[1] - slope_range[0])
  if inverted:
    slope = -slope
  bp, npx, npy, nix, niy = _hf_inner(size, horizontal_scale, border_width)
  span = nix * horizontal_scale if bp > 0 else size[0]
  hmax = int(slope * span / 2 / vertical_scale)
  cx, cy = int(nix / 2), int(niy / 2)
  tx = ((cx - np.abs(cx - np.arange(nix))) / cx).reshape(nix, 1)
  ty = ((cy - np.abs(cy - np.arange(niy))) / cy).reshape(1, niy)
  raw = hmax * tx * ty
  half_pf = int(platform_width / horizontal_scale / 2)
  ix, iy = nix // 2 - half_pf, niy // 2 - half_pf
  zpf = raw[ix, iy] if (bp == 0 or (ix >= 0 and iy >= 0)) else 0
  raw = np.clip(raw, min(0, zpf), max(0, zpf))
  levels = np.zeros((npx, npy), dtype=np.int16)
  levels[bp:npx - bp, bp:npy - bp] = np.rint(raw).astype(np.int16)
  data, hsize, elev = _hf_finish(levels, size, vertical_scale, base_thickness_ratio, 0.0)
  z = -elev if inverted else 0.0
  spawn = (z if inverted else elev) - hsize[3]
  return HeightField(data, hsize, np.array([size[0] / 2, size[1] / 2, z])), np.array([size[0] / 2, size[1] / 2, spawn])


def hf_random_uniform(size, rng, noise_range, noise_step=0.005, downsampled_scale=None, horizontal_scale=0.1,
                      vertical_scale=0.005, base_thickness_ratio=1.0, border_width=0.0):
  """Random levels on a coarse grid, bicubic-spline interpolated to the sample grid
  (``heightfield_terrains.py:255-391``)."""
  from scipy.interpolate import RectBivariateSpline

  ds = horizontal_scale if downsampled_scale is None else downsampled_scale
  if ds < horizontal_scale:
    raise ValueError(f"Downsampled scale must be >= horizontal scale: {ds} < {horizontal_scale}")
  bp, npx, npy, nix, niy = _hf_inner(size, horizontal_scale, border_width)
  ext = (nix * horizontal_scale, niy * horizontal_scale) if bp > 0 else (size[0], size[1])
  ncx, ncy = int(ext[0] / ds), int(ext[1] / ds)
  lv = [int(v / vertical_scale) for v in (noise_range[0], noise_range[1], noise_step)]
  coarse = rng.choice(np.arange(lv[0], lv[1] + lv[2], lv[2]), size=(ncx, ncy))
  f = RectBivariateSpline(np.linspace(0, ext[0], ncx), np.linspace(0, ext[1], ncy), coarse)
  fine = f(np.linspace(0, ext[0], nix), np.linspace(0, ext[1], niy))
  levels = np.zeros((npx, npy), dtype=np.int16)
  levels[bp:npx - bp, bp:npy - bp] = np.rint(fine).astype(np.int16)
  data, hsize, _ = _hf_finish(levels, size, vertical_scale, base_thickness_ratio, 0.0)
  origin = np.array([size[0] / 2, size[1] / 2, (noise_range[0] + noise_range[1]) / 2])
  return HeightField(data, hsize, np.array([size[0] / 2, size[1] / 2, 0.0])), origin


def hf_wave(size, difficulty, amplitude_range, num_waves=1.0, horizontal_scale=0.1, vertical_scale=0.005,
            base_thickness_ratio=0.25, border_width=0.0):
  """cos(y) + sin(x) waves, centred on z = 0 (``heightfield_terrains.py:394-499``)."""
  if num_waves <= 0:
    raise ValueError(f"Number of waves must be positive. Got: {num_waves}")
  amp = amplitude_range[0] + difficulty * (amplitude_range[1] - amplitude_range[0])
  bp, npx, npy, nix, niy = _hf_inner(size, horizontal_scale, border_width)
  ap = int(0.5 * amp / vertical_scale)
  k = 2 * np.pi / (niy / num_waves)
  raw = ap * (np.cos(np.arange(niy) * k).reshape(1, niy) + np.sin(np.arange(nix) * k).reshape(nix, 1))
  levels = np.zeros((npx, npy), dtype=np.int16)
  levels[bp:npx - bp, bp:npy - bp] = np.rint(raw).astype(np.int16)
  data, hsize, elev = _hf_finish(levels, size, vertical_scale, base_thickness_ratio, 0.0)
  return (HeightField(data, hsize, np.array([size[0] / 2, size[1] / 2, -elev / 2])),
          np.array([size[0] / 2, size[1] / 2, 0.0]))


@dataclass
class RoughTerrainCfg:
  """``ROUGH_TERRAINS_CFG`` (``terrains/config.py:7-27``) with ``curriculum=True`` as set by the velocity
  task (``tasks/velocity/velocity_env_cfg.py:274-278``)."""

  size: tuple = (8.0, 8.0)
  border_width: float = 20.0
  border_height: float = 1.0
  num_rows: int = 10
  num_cols: int = 20
  curriculum: bool = True
  seed: int = 0
  difficulty_range: tuple = (0.0, 1.0)
  # The reference's pyramid height fields report a spawn origin at the level of the field's base
  # (heightfield_terrains.py:245-249: max_height - base_thickness, i.e. inside the solid for the default ratio 1);
  # they are disabled upstream, so nothing depends on it.  True: spawn on the surface at the patch centre instead.
  hf_spawn_on_surface: bool = True
  # (kind, proportion, kwargs) in insertion order
  sub_terrains: tuple = field(default_factory=lambda: (
    ("flat", 0.4, {}),
    ("pyramid_stairs", 0.3, dict(step_height_range=(0.0, 0.1), step_width=0.3, platform_width=3.0, border_width=1.0)),
    ("pyramid_stairs_inv", 0.3, dict(step_height_range=(0.0, 0.1), step_width=0.3, platform_width=3.0, border_width=1.0)),
  ))


_HF = dict(border_width=0.25)
# ROUGH_TERRAINS_CFG with the four height-field entries the reference keeps commented out (terrains/config.py:28-55):
# same names, parameters and proportions (the generator normalises the sum).
FULL_SUB_TERRAINS = RoughTerrainCfg().sub_terrains + (
  ("hf_pyramid_slope", 0.1, dict(slope_range=(0.0, 1.0), platform_width=2.0, **_HF)),
  ("hf_pyramid_slope_inv", 0.1, dict(slope_range=(0.0, 1.0), platform_width=2.0, **_HF)),
  ("random_rough", 0.2, dict(noise_range=(0.02, 0.10), noise_step=0.02, **_HF)),
  ("wave_terrain", 0.2, dict(amplitude_range=(0.0, 0.2), num_waves=4, **_HF)),
)


def _make(kind, size, difficulty, kw, rng=None):
  if kind == "flat":
    return flat_patch(size)
  if kind == "pyramid_stairs":
    return pyramid_stairs(size, difficulty, inverted=False, **kw)
  if kind == "pyramid_stairs_inv":
    return pyramid_stairs(size, difficulty, inverted=True, **kw)
  # height-field sub-terrains (commented out of the reference's ROUGH_TERRAINS_CFG, terrains/config.py:28-50, because
  # mujoco-warp could not compile them; here they collide through csrc/b2_convex.h)
  if kind in ("hf_pyramid_slope", "hf_pyramid_slope_inv"):
    hf, org = hf_pyramid_sloped(size, difficulty, inverted=kind.endswith("_inv"), **kw)
    return [hf], org
  if kind == "random_rough":
    hf, org = hf_random_uniform(size, rng, **kw)
    return [hf], org
  if kind == "wave_terrain":
    hf, org = hf_wave(size, difficulty, **kw)
    return [hf], org
  raise NotImplementedError(f"sub-terrain '{kind}'")


def generate_terrain(cfg: RoughTerrainCfg):
  """Returns (boxes in geom order, terrain_origins[num_rows, num_cols, 3])."""
  rng = np.random.default_rng(cfg.seed)
  prop = np.array([p for _, p, _ in cfg.sub_terrains], dtype=float)
  prop /= prop.sum()
  origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
  boxes: list[Box] = []

  def corner(r, c):
    return np.array([-cfg.num_rows * cfg.size[0] / 2 + r * cfg.size[0],
                     -cfg.num_cols * cfg.size[1] / 2 + c * cfg.size[1], 0.0])

  def place(r, c, kind, kw, difficulty):
    bx, org = _make(kind, cfg.size, difficulty, kw, rng)
    w = corner(r, c)
    for it in bx:
      boxes.append(HeightField(it.data, it.size, it.pos + w) if isinstance(it, HeightField) else (it[0], it[1] + w))
      if isinstance(it, HeightField) and cfg.hf_spawn_on_surface:
        org = org.copy()
        org[2] = it.pos[2] + it.data[it.data.shape[0] // 2, it.data.shape[1] // 2] * it.size[2]
    origins[r, c] = org + w

  lo, hi = cfg.difficulty_range
  if cfg.curriculum:  # type by column, difficulty by row (terrain_generator.py:149-175)
    cum = np.cumsum(prop)
    col_kind = [int(np.min(np.where(c / cfg.num_cols + 0.001 < cum)[0])) for c in range(cfg.num_cols)]
    for c in range(cfg.num_cols):
      for r in range(cfg.num_rows):
        d = lo + (hi - lo) * (r + rng.uniform()) / cfg.num_rows
        kind, _, kw = cfg.sub_terrains[col_kind[c]]
        place(r, c, kind, kw, d)
  else:  # random type and difficulty per patch (:121-147)
    for idx in range(cfg.num_rows * cfg.num_cols):
      r, c = divmod(idx, cfg.num_cols)
      k = int(rng.choice(len(prop), p=prop))
      kind, _, kw = cfg.sub_terrains[k]
      place(r, c, kind, kw, rng.uniform(lo, hi))
  inner = (cfg.num_rows * cfg.size[0], cfg.num_cols * cfg.size[1])
  outer = (inner[0] + 2 * cfg.border_width, inner[1] + 2 * cfg.border_width)
  if cfg.border_width > 0:
    boxes += ring_boxes(outer, inner, abs(cfg.border_height), (0.0, 0.0, -cfg.border_height / 2))
  return boxes, origins


def terrain_spec(cfg: RoughTerrainCfg, name: str = "terrain"):
  """Spec with one static body ``terrain`` holding the boxes as geoms ``terrain_<i>``."""
  boxes, origins = generate_terrain(cfg)
  sp = S.Spec()
  body = sp.worldbody.add_body(name=name)
  for i, it in enumerate(boxes):
    if isinstance(it, HeightField):
      hf = sp.add_hfield(f"hfield_{i}", size=it.size, nrow=it.data.shape[0], ncol=it.data.shape[1], userdata=it.data)
      body.add_geom(name=f"{name}_{i}", type=S.GEOM_HFIELD, hfieldname=hf.name, pos=tuple(it.pos))
    else:
      body.add_geom(name=f"{name}_{i}", type=S.GEOM_BOX, size=tuple(it[0]), pos=tuple(it[1]))
  return sp, origins


def env_origins_curriculum(num_envs: int, origins: np.ndarray, max_init_level: int = 5, seed: int = 0):
  """Spawn levels / types per env as ``terrain_importer.py:203-223``: random row <= max level, columns
  assigned round-robin."""
  rng = np.random.default_rng(seed)
  rows, cols = origins.shape[:2]
  level = rng.integers(0, min(max_init_level, rows - 1) + 1, size=num_envs)
  kind = np.floor(np.arange(num_envs) / (num_envs / cols)).astype(int) % cols
  return origins[level, kind], level, kind
