"""``warp`` stand-in: the device / stream / graph / array plumbing that mjlab's sim layer calls
(``src/mjlab/__init__.py:4-19``, ``sim/sim.py:101-195``, ``sim/sim_data.py:15-229``, ``sim/randomization.py:9-55``),
implemented over torch.  A ``wp.array`` here is a thin view of a ``torch.Tensor`` (engine memory or torch memory).
"""

from __future__ import annotations

import threading
from types import SimpleNamespace
from typing import Any

import torch

__b2_compat__ = True
__version__ = "0.0-b2compat"

config = SimpleNamespace(enable_backward=False, quiet=True, verbose=False, version="b2sim-standin")  # (manager_based_rl_env.py:38 reads wp.config.version)

float32 = torch.float32
int32 = torch.int32
vec3 = torch.float32
Any_ = Any


class Device:
  def __init__(self, alias: str):
    d = torch.device(alias)
    self.alias = str(d)
    self.is_cuda = d.type == "cuda"
    self.is_cpu = d.type == "cpu"
    self.ordinal = d.index if d.index is not None else 0
    self.torch_device = d

  def __repr__(self):
    return f"wp.Device({self.alias})"

  def __eq__(self, o):
    return isinstance(o, Device) and o.alias == self.alias

  def __hash__(self):
    return hash(self.alias)


def get_device(alias=None) -> Device:
  if isinstance(alias, Device):
    return alias
  if alias is None:
    alias = "cuda:0" if torch.cuda.is_available() else "cpu"
  return Device(str(alias))


def is_mempool_enabled(device) -> bool:
  """mjlab captures CUDA graphs when this is true (``sim/sim.py:121-123``); torch's caching allocator plays
  the role of Warp's mempool, so graphs are available on every CUDA device."""
  return get_device(device).is_cuda


class ScopedDevice:
  def __init__(self, device):
    self.device = get_device(device)
    self._ctx = None

  def __enter__(self):
    if self.device.is_cuda:
      self._ctx = torch.cuda.device(self.device.torch_device)
      self._ctx.__enter__()
    return self.device

  def __exit__(self, *exc):
    if self._ctx is not None:
      self._ctx.__exit__(*exc)
    return False


class Stream:
  def __init__(self, device):
    self.device = get_device(device)

  @property
  def cuda_stream(self) -> int:
    return torch.cuda.current_stream(self.device.torch_device).cuda_stream


def get_stream(device=None) -> Stream:
  return Stream(device)


class ScopedStream:
  def __init__(self, stream):
    self.stream = stream

  def __enter__(self):
    return self.stream

  def __exit__(self, *exc):
    return False


class ScopedCapture:
  """``with wp.ScopedCapture() as capture: ...; capture.graph`` over ``torch.cuda.CUDAGraph``.  Work issued on
  torch's current stream inside the block (the engine launches there) is captured."""

  def __init__(self, device=None, **_):
    self.graph = None
    self._side = None
    self._cm = None

  def __enter__(self):
    dev = torch.cuda.current_device()
    self._side = torch.cuda.Stream(dev)
    self._side.wait_stream(torch.cuda.current_stream(dev))
    self.graph = torch.cuda.CUDAGraph()
    self._cm = torch.cuda.graph(self.graph, stream=self._side)
    self._cm.__enter__()
    return self

  def __exit__(self, *exc):
    r = self._cm.__exit__(*exc)
    torch.cuda.current_stream().wait_stream(self._side)
    return r


def capture_launch(graph) -> None:
  graph.replay()


def synchronize() -> None:
  if torch.cuda.is_available():
    torch.cuda.synchronize()


class array:
  """View of a torch tensor with the attributes mjlab reads from ``wp.array``: ``shape``, ``strides`` (bytes),
  ``dtype``, ``device``, ``ptr``, slicing, ``flatten()``, ``numpy()``."""

  def __init__(self, data=None, dtype=None, shape=None, device=None, ndim=None, **_):
    if data is None and shape is None:
      self._t = None  # `wp.array(dtype=...)` used as a type annotation in a kernel signature
      return
    if isinstance(data, torch.Tensor):
      t = data
    elif data is not None:
      t = torch.as_tensor(data, dtype=dtype if isinstance(dtype, torch.dtype) else None,
                          device=get_device(device).torch_device if device is not None else None)
    else:
      if isinstance(shape, int):
        shape = (shape,)
      t = torch.empty(tuple(shape), dtype=dtype if isinstance(dtype, torch.dtype) else torch.float32,
                      device=get_device(device).torch_device if device is not None else None)
    self._t = t

  # -- attributes ------------------------------------------------------------------------------------
  @property
  def shape(self):
    return tuple(self._t.shape)

  @property
  def strides(self):
    return tuple(s * self._t.element_size() for s in self._t.stride())

  @property
  def dtype(self):
    return self._t.dtype

  @property
  def ndim(self):
    return self._t.dim()

  @property
  def device(self) -> Device:
    return Device(str(self._t.device))

  @property
  def ptr(self) -> int:
    return self._t.data_ptr()

  @property
  def size(self) -> int:
    return self._t.numel()

  def __len__(self):
    return self._t.shape[0]

  def __repr__(self):
    return f"wp.array(shape={self.shape}, dtype={self.dtype}, device={self._t.device})"

  # -- views -----------------------------------------------------------------------------------------
  def __getitem__(self, idx):
    if isinstance(idx, torch.Tensor):  # element-wise kernel body: gather
      return self._t[idx.long()]
    return array(self._t[idx])

  def __setitem__(self, idx, value):
    if isinstance(idx, torch.Tensor):  # element-wise kernel body: scatter
      self._t[idx.long()] = value._t if isinstance(value, array) else value
    else:
      self._t[idx] = value._t if isinstance(value, array) else value

  def flatten(self) -> "array":
    t = self._t
    if t.dim() > 0 and t.stride(0) == 0 and t.shape[0] > 1:
      t = t[:1]  # one copy shared by all worlds: its flat form is that copy
    return array(t.reshape(-1))

  def reshape(self, *shape):
    return array(self._t.reshape(*shape))

  def numpy(self):
    return self._t.detach().cpu().numpy()

  def zero_(self):
    self._t.zero_()
    return self

  def fill_(self, v):
    self._t.fill_(v)
    return self


def _typed_array(nd):
  def make(data=None, dtype=None, shape=None, device=None, **kw):
    return array(data, dtype=dtype, shape=shape, device=device, **kw)

  make.__name__ = f"array{nd}d"
  return make


array1d, array2d, array3d, array4d = (_typed_array(k) for k in (1, 2, 3, 4))


def to_torch(a, requires_grad=None):
  return a._t if isinstance(a, array) else a


def from_torch(t, dtype=None, **_):
  return array(t)


def zeros(shape, dtype=torch.float32, device=None, **_):
  return array(None, dtype=dtype, shape=shape, device=device).zero_()


def empty(shape, dtype=torch.float32, device=None, **_):
  return array(None, dtype=dtype, shape=shape, device=device)


# ---- element-wise kernels ---------------------------------------------------------------------------
_tls = threading.local()


class Kernel:
  def __init__(self, fn):
    self.fn = fn
    self.__name__ = getattr(fn, "__name__", "kernel")


def kernel(fn=None, **_):
  """``@wp.kernel`` / ``@wp.kernel(module="unique")``: the function body is kept and run vectorised by
  ``launch`` (one tensor of thread ids instead of one thread per id)."""
  if fn is None:
    return lambda f: Kernel(f)
  return Kernel(fn)


def func(fn=None, **_):
  return fn if fn is not None else (lambda f: f)


def tid():
  return _tls.tid


def launch(kernel, dim, inputs=(), outputs=(), device=None, **_):
  n = dim if isinstance(dim, int) else int(dim[0])
  args = list(inputs) + list(outputs)
  dev = None
  for a in args:
    if isinstance(a, array):
      dev = a._t.device
      break
  _tls.tid = torch.arange(n, device=dev)
  try:
    kernel.fn(*args)
  finally:
    _tls.tid = None
