"""ctypes binding of libb2sim.so (the C ABI declared in include/b2sim.h)."""

from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np

from mjlab_b200.compiler.compile import MODEL_ARRAYS, MODEL_SCALARS_F, MODEL_SCALARS_I, fill_asset_defaults

import os

LIB_PATH = Path(os.environ.get("B2SIM_LIB") or Path(__file__).resolve().parents[1] / "csrc" / "libb2sim.so")


class B2Array(ctypes.Structure):
  _fields_ = [
    ("name", ctypes.c_char_p), ("dtype", ctypes.c_int32), ("n", ctypes.c_int64),
    ("data", ctypes.c_void_p),
  ]


class B2ModelDesc(ctypes.Structure):
  _fields_ = [
    ("narray", ctypes.c_int32), ("arrays", ctypes.POINTER(B2Array)),
    ("gravity", ctypes.c_double * 3),
  ]


class B2Tensor(ctypes.Structure):
  _fields_ = [
    ("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32),
    ("shape", ctypes.c_int64 * 4), ("stride", ctypes.c_int64 * 4), ("device", ctypes.c_int32),
  ]


class B2Stats(ctypes.Structure):
  _fields_ = [
    ("ncon_max", ctypes.c_int32), ("ncon_cap", ctypes.c_int32), ("nefc_max", ctypes.c_int32),
    ("nefc_cap", ctypes.c_int32), ("overflow_worlds", ctypes.c_int32),
    ("niter_max", ctypes.c_int32), ("ncon_mean", ctypes.c_double),
    ("nefc_mean", ctypes.c_double), ("niter_mean", ctypes.c_double),
  ]


class B2VelEnvArgs(ctypes.Structure):
  _fields_ = (
    [(n, ctypes.c_void_p) for n in (
      "action", "U", "default_qpos", "default_joint_pos", "action_scale", "soft_lo", "soft_hi",
      "env_origins", "episode_length", "last_action", "command", "push_time_left", "obs", "reward",
      "terminated", "truncated", "done", "cmd_time_left", "heading_target", "is_standing", "critic", "log_row")]
    + [(n, ctypes.c_float) for n in ("step_dt", "fall_angle", "push_vel", "push_lo", "push_hi")]
    + [("max_episode_length", ctypes.c_int32)]
  )


class B2TrackEnvArgs(ctypes.Structure):
  _fields_ = (
    [(n, ctypes.c_void_p) for n in (
      "action", "U", "default_joint_pos", "soft_lo", "soft_hi", "env_origins", "m_joint_pos", "m_joint_vel",
      "m_body_pos", "m_body_quat", "m_body_lin", "m_body_ang", "body_idx", "ee_idx", "time_steps",
      "episode_length", "last_action", "push_time_left", "body_pos_rel", "body_quat_rel", "reward", "terminated",
      "truncated", "mask", "log_row", "obs", "critic")]
    + [("pose_range", ctypes.c_float * 12), ("vel_range", ctypes.c_float * 12)]
    + [(n, ctypes.c_float) for n in ("step_dt", "jp_lo", "jp_hi", "push_lo", "push_hi")]
    + [(n, ctypes.c_int32) for n in ("nb", "nee", "anchor", "T", "self_collision_adr", "root_body", "max_episode_length")]
  )


def make_model_desc(model):
  """Pack a compiled Model into a B2ModelDesc. Returns (desc, keepalive)."""
  keep = []
  names = [n for n, _ in MODEL_ARRAYS] + MODEL_SCALARS_I + MODEL_SCALARS_F
  fill_asset_defaults(model.arrays)
  arrs = (B2Array * len(names))()
  for i, n in enumerate(names):
    a = np.asarray(model.arrays[n])
    if a.dtype.kind in "iu":
      a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
      dt = 1
    else:
      a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
      dt = 0
    keep.append(a)
    nb = n.encode()
    keep.append(nb)
    arrs[i].name = nb
    arrs[i].dtype = dt
    arrs[i].n = a.size
    arrs[i].data = a.ctypes.data if a.size else None
  desc = B2ModelDesc()
  desc.narray = len(names)
  desc.arrays = arrs
  for k in range(3):
    desc.gravity[k] = float(model.opt_gravity[k])
  keep.append(arrs)
  return desc, keep


_lib = None


def load_library() -> ctypes.CDLL:
  """Load libb2sim.so; fails loudly when the CUDA extension has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not LIB_PATH.exists():
    raise RuntimeError(
      f"{LIB_PATH} not found: build the CUDA extension first "
      "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
    )
  L = ctypes.CDLL(str(LIB_PATH))
  L.b2_last_error.restype = ctypes.c_char_p
  L.b2_version.restype = ctypes.c_char_p
  L.b2_field_name.restype = ctypes.c_char_p
  L.b2_launch_count.restype = ctypes.c_int64
  vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
  L.b2_create.argtypes = [ctypes.POINTER(B2ModelDesc), ci, ci, ci, ci, ctypes.POINTER(vp)]
  L.b2_destroy.argtypes = [vp]
  L.b2_get_field.argtypes = [vp, ci, ctypes.c_char_p, ctypes.POINTER(B2Tensor)]
  L.b2_num_fields.argtypes = [vp, ci]
  L.b2_field_name.argtypes = [vp, ci, ci]
  L.b2_expand_model_field.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(B2Tensor)]
  L.b2_set_option.argtypes = [vp, ctypes.c_char_p, cd]
  L.b2_get_option.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(cd)]
  L.b2_step.argtypes = [vp, vp]
  L.b2_forward.argtypes = [vp, vp]
  L.b2_step_n.argtypes = [vp, ci, vp]
  L.b2_forward_masked.argtypes = [vp, vp, vp]
  L.b2_velenv_pre.argtypes = [vp, vp, vp, vp, vp]
  L.b2_velenv_post.argtypes = [vp, ctypes.POINTER(B2VelEnvArgs), vp]
  L.b2_trackenv_post1.argtypes = [vp, ctypes.POINTER(B2TrackEnvArgs), vp]
  L.b2_trackenv_post2.argtypes = [vp, ctypes.POINTER(B2TrackEnvArgs), vp]
  L.b2_step_host.argtypes = [vp, vp, ci, vp, vp, vp]
  L.b2_stats.argtypes = [vp, vp, ctypes.POINTER(B2Stats)]
  L.b2_launch_count.argtypes = [vp]
  L.b2_algorithmic_bytes.argtypes = [vp, vp, ctypes.POINTER(cd), ctypes.POINTER(cd)]
  _lib = L
  return L


def check(rc: int) -> None:
  if rc != 0:
    raise RuntimeError(f"b2sim: {load_library().b2_last_error().decode()}")
