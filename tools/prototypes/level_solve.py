"""Prototype for DESIGN.md §9.1: level-scheduled triangular solves for the bottom-up L^T D L factor.

numpy only (no GPU): builds the dof-tree level sets of a compiled model, runs the two sweeps level by level
(all pivots of one tree depth are independent, so a warp can issue their shuffles together) and checks the result
against a dense solve.  Prints the dependency depth a level schedule needs versus the 2 x nv steps of the
sequential sweeps in `ldl_solve` (mjlab_b200/csrc/b2_kernel.cuh).

  python tools/prototypes/level_solve.py g1_flat go1_flat
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from mjlab_b200.asset_zoo import load_compiled  # noqa: E402


def levels_of(parent):
  n = len(parent)
  depth = np.zeros(n, dtype=int)
  for k in range(n):
    depth[k] = 0 if parent[k] < 0 else depth[parent[k]] + 1
  return [np.nonzero(depth == d)[0] for d in range(depth.max() + 1)], depth


def ltdl(M, parent):
  """Bottom-up L^T D L restricted to the tree pattern (mj_factorI order): returns unit-lower L and d."""
  n = len(M)
  A = M.copy()
  for k in range(n - 1, -1, -1):
    i = parent[k]
    while i >= 0:
      t = A[k, i] / A[k, k]
      j = i
      while j >= 0:
        A[i, j] -= t * A[k, j]
        j = parent[j]
      A[k, i] = t
      i = parent[i]
  L = np.tril(A, -1) + np.eye(n)
  return L, np.diag(A).copy()


def solve_by_levels(L, d, b, lv):
  x = b.copy()
  for s in reversed(lv):            # L^T y = b: deepest level first; pivots of one level never touch each other
    for k in s:
      x -= L[k] * x[k] * (np.arange(len(x)) < k)
  x /= d
  for s in lv:                      # L x = z: root level first
    for k in s:
      x[k] -= L[k, :k] @ x[:k]
  return x


def main():
  for name in sys.argv[1:] or ["g1_flat", "go1_flat"]:
    m = load_compiled(name)
    par = [int(p) for p in m.dof_parentid]
    n = len(par)
    lv, depth = levels_of(par)
    rng = np.random.default_rng(0)
    M = np.zeros((n, n))
    for k in range(n):              # random SPD matrix with the tree's sparsity
      idx, p = [k], par[k]
      while p >= 0:
        idx.append(p)
        p = par[p]
      v = np.zeros(n)
      v[idx] = rng.normal(size=len(idx))
      M += np.outer(v, v)
    M += np.diag(rng.uniform(0.1, 1.0, n))
    L, d = ltdl(M, par)
    assert np.abs(L.T @ np.diag(d) @ L - M).max() < 1e-10
    b = rng.normal(size=n)
    x = solve_by_levels(L, d, b, lv)
    err = np.abs(M @ x - b).max()
    widths = [len(s) for s in lv]
    print(f"{name}: nv={n}  tree levels={len(lv)} (widths {widths})  sequential sweep steps={2 * n}  "
          f"level steps={2 * len(lv)}  residual={err:.1e}")


if __name__ == "__main__":
  main()
