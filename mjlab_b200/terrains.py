"""Procedural box terrains (host-side scene setup for BASELINE config E: Go1 rough terrain).

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
  # (kind, proportion, kwargs) in insertion order
  sub_terrains: tuple = field(default_factory=lambda: (
    ("flat", 0.4, {}),
    ("pyramid_stairs", 0.3, dict(step_height_range=(0.0, 0.1), step_width=0.3, platform_width=3.0, border_width=1.0)),
    ("pyramid_stairs_inv", 0.3, dict(step_height_range=(0.0, 0.1), step_width=0.3, platform_width=3.0, border_width=1.0)),
  ))


def _make(kind, size, difficulty, kw):
  if kind == "flat":
    return flat_patch(size)
  if kind == "pyramid_stairs":
    return pyramid_stairs(size, difficulty, inverted=False, **kw)
  if kind == "pyramid_stairs_inv":
    return pyramid_stairs(size, difficulty, inverted=True, **kw)
  raise NotImplementedError(f"sub-terrain '{kind}' (heightfields are disabled upstream too, terrains/config.py:28-29)")


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
    bx, org = _make(kind, cfg.size, difficulty, kw)
    w = corner(r, c)
    boxes.extend((h, p + w) for h, p in bx)
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
  for i, (half, pos) in enumerate(boxes):
    body.add_geom(name=f"{name}_{i}", type=S.GEOM_BOX, size=tuple(half), pos=tuple(pos))
  return sp, origins


def env_origins_curriculum(num_envs: int, origins: np.ndarray, max_init_level: int = 5, seed: int = 0):
  """Spawn levels / types per env as ``terrain_importer.py:203-223``: random row <= max level, columns
  assigned round-robin."""
  rng = np.random.default_rng(seed)
  rows, cols = origins.shape[:2]
  level = rng.integers(0, min(max_init_level, rows - 1) + 1, size=num_envs)
  kind = np.floor(np.arange(num_envs) / (num_envs / cols)).astype(int) % cols
  return origins[level, kind], level, kind
