"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE — see oracle/b2_oracle.c header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
"""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from mjlab_b200.sim.native import make_model_desc

_DIR = Path(__file__).parent


def build(force: bool = False) -> None:
  if force or not (_DIR / "libb2oracle64.so").exists() or not (_DIR / "libb2oracle32.so").exists():
    subprocess.check_call(["make", "-C", str(_DIR), "-s"] + (["-B"] if force else []))


class Oracle:
  """One batched oracle instance. Fields are numpy views ``[nworld, n]`` of its memory."""

  def __init__(self, model, nworld: int = 1, maxcon: int = 64, njmax: int = 300,
               precision: str = "f64"):
    build()
    self.lib = ctypes.CDLL(str(_DIR / f"libb2oracle{64 if precision == 'f64' else 32}.so"))
    self.dtype = np.float64 if precision == "f64" else np.float32
    L = self.lib
    L.b2o_create.restype = ctypes.c_void_p
    L.b2o_field.restype = ctypes.c_void_p
    L.b2o_ifield.restype = ctypes.c_void_p
    L.b2o_model_field.restype = ctypes.c_void_p
    self.model = model
    self.nworld = nworld
    desc, self._keep = make_model_desc(model)
    self.h = ctypes.c_void_p(
      L.b2o_create(ctypes.byref(desc), ctypes.c_int(nworld), ctypes.c_int(maxcon),
                   ctypes.c_int(njmax))
    )
    self._cache = {}

  def __del__(self):
    try:
      self.lib.b2o_destroy(self.h)
    except Exception:
      pass

  def field(self, name: str) -> np.ndarray:
    if name in self._cache:
      return self._cache[name]
    n = ctypes.c_int(0)
    p = self.lib.b2o_field(self.h, name.encode(), ctypes.byref(n))
    if p:
      ct = ctypes.c_double if self.dtype == np.float64 else ctypes.c_float
      arr = np.ctypeslib.as_array((ct * (self.nworld * n.value)).from_address(p))
    else:
      p = self.lib.b2o_ifield(self.h, name.encode(), ctypes.byref(n))
      if not p:
        raise KeyError(name)
      arr = np.ctypeslib.as_array((ctypes.c_int * (self.nworld * n.value)).from_address(p))
    arr = arr.reshape(self.nworld, n.value)
    self._cache[name] = arr
    return arr

  def __getattr__(self, name):
    if name.startswith("_") or name in ("lib", "h", "model", "nworld", "dtype"):
      raise AttributeError(name)
    return self.field(name)

  def model_field(self, name: str) -> np.ndarray:
    n = ctypes.c_int(0)
    p = self.lib.b2o_model_field(self.h, name.encode(), ctypes.byref(n))
    if not p:
      raise KeyError(name)
    ct = ctypes.c_double if self.dtype == np.float64 else ctypes.c_float
    return np.ctypeslib.as_array((ct * (self.nworld * n.value)).from_address(p)).reshape(
      self.nworld, n.value
    )

  def set_option(self, key: str, value: float) -> None:
    self.lib.b2o_set_option(self.h, key.encode(), ctypes.c_double(value))

  def forward(self, nthread: int = 0) -> None:
    self.lib.b2o_forward(self.h, ctypes.c_int(nthread))

  def step(self, nthread: int = 0) -> None:
    self.lib.b2o_step(self.h, ctypes.c_int(nthread))

  def max_threads(self) -> int:
    return int(self.lib.b2o_max_threads())
