"""Minimal stand-in for the ``gymnasium`` wheel (absent from this image): the base classes, spaces and registry
calls that ``mjlab.envs`` / ``mjlab.tasks`` make at import and construction time.  Test infrastructure
(tests/ref_runner.py puts tests/stubs on the path); nothing here is on the product path."""
import types

import numpy as np


class Space:
  def __init__(self, shape=None, dtype=None):
    self.shape, self.dtype = shape, dtype


class Box(Space):
  def __init__(self, low=-np.inf, high=np.inf, shape=None, dtype=np.float32):
    super().__init__(tuple(shape) if shape is not None else np.shape(low), dtype)
    self.low, self.high = low, high


class Dict(Space, dict):
  def __init__(self, spaces=None, **kw):
    Space.__init__(self)
    dict.__init__(self, spaces or {}, **kw)

  @property
  def spaces(self):
    return self


spaces = types.ModuleType("gymnasium.spaces")
spaces.Space, spaces.Box, spaces.Dict = Space, Box, Dict


class Env:
  metadata: dict = {}
  observation_space = None
  action_space = None

  @property
  def unwrapped(self):
    return self

  def close(self):
    pass


class Wrapper(Env):
  def __init__(self, env):
    self.env = env

  def __getattr__(self, name):
    return getattr(self.env, name)

  @property
  def unwrapped(self):
    return self.env.unwrapped


class _Spec:
  def __init__(self, id, entry_point=None, kwargs=None, **extra):
    self.id, self.entry_point, self.kwargs = id, entry_point, dict(kwargs or {})
    self.__dict__.update(extra)


registry: dict = {}


def register(id, entry_point=None, kwargs=None, **extra):
  registry[id] = _Spec(id, entry_point, kwargs, **extra)


def spec(id):
  return registry[id]


def make(id, **kw):
  import importlib

  s = registry[id]
  ep = s.entry_point
  if isinstance(ep, str):
    mod, _, attr = ep.partition(":")
    ep = getattr(importlib.import_module(mod), attr)
  return ep(**{**s.kwargs, **kw})


def _batch_space(space, n=1):
  if isinstance(space, Dict):
    return Dict({k: _batch_space(v, n) for k, v in space.items()})
  return Box(low=getattr(space, "low", -np.inf), high=getattr(space, "high", np.inf), shape=(n, *(space.shape or ())),
             dtype=space.dtype or np.float32)


vector = types.ModuleType("gymnasium.vector")
vector.utils = types.ModuleType("gymnasium.vector.utils")
vector.utils.batch_space = _batch_space
wrappers = types.ModuleType("gymnasium.wrappers")
wrappers.RecordVideo = Wrapper

import sys as _sys  # noqa: E402

for _m in (spaces, vector, vector.utils, wrappers):
  _sys.modules[_m.__name__] = _m
