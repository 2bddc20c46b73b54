"""Sobol sequences in natural order, generated on the device.

Mirrors `math/random_ops/sobol/sobol_impl.py`: `sample` (39-167) and the
direction numbers (`_compute_direction_numbers` 171-197, `load_data` 237-261),
which are computed once by libtqf and cached (the reference rebuilds them in a
Python triple loop on every call).
"""
import ctypes as C
import os
import threading

import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor

_MAX_POSITIVE = 2**31 - 1
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..',
                     '..', 'data', 'joe_kuo_6_21201.npz')
_lock = threading.Lock()
_table = None
_direction_cache = {}


def _joe_kuo():
  global _table
  with _lock:
    if _table is None:
      with np.load(_DATA) as z:
        _table = (np.ascontiguousarray(z['a'], dtype=np.uint32),
                  np.ascontiguousarray(z['s'], dtype=np.uint8),
                  np.ascontiguousarray(z['m'], dtype=np.uint32))
    return _table


def direction_numbers(dim):
  """int32 [dim, 32] matrix of the m_{k,j} (cached per dim)."""
  dim = int(dim)
  with _lock:
    hit = _direction_cache.get(dim)
  if hit is not None:
    return hit
  if dim < 1 or dim > 21201:
    raise ValueError('Sobol dimension must be in [1, 21201], got {}'.format(dim))
  a, s, m = _joe_kuo()
  out = np.empty((dim, 32), dtype=np.int32)
  _lib.check(_lib.lib().tqf_sobol_direction_numbers(
      a.ctypes.data, s.ctypes.data, m.ctypes.data, a.shape[0], dim,
      out.ctypes.data))
  with _lock:
    _direction_cache[dim] = out
  return out


def direction_numbers_from_file(path, dim):
  """Same, parsing an original `new-joe-kuo-6.21201` text file."""
  out = np.empty((int(dim), 32), dtype=np.int32)
  _lib.check(_lib.lib().tqf_sobol_direction_numbers_from_file(
      os.fsencode(path), int(dim), out.ctypes.data))
  return out


def _fill(dim, num_results, skip, kind, dtype, first_result=0, count=None):
  dtype = _tensor.np_dtype(dtype)
  count = num_results if count is None else count
  out_dtype = np.int32 if kind == 0 else dtype
  out = _tensor.empty((count, dim), out_dtype)
  dn = direction_numbers(dim)
  _lib.check(_lib.lib().tqf_sobol_fill(
      dn.ctypes.data, dim, num_results, skip, first_result, count, kind,
      _tensor.tqf_dtype(dtype), out.data_ptr(), _tensor.current_stream_ptr()))
  return out


def sample(dim, num_results, skip=0, validate_args=False, dtype=None,
           name=None):
  """`sobol.sample`: `[num_results, dim]` points in (0, 1), natural order."""
  del name
  dim, num_results, skip = int(dim), int(num_results), int(skip)
  if validate_args:
    if dim <= 0:
      raise ValueError('`dim` must be greater than zero')
    if num_results <= 0:
      raise ValueError('`num_results` must be greater than zero')
    if skip < 0:
      raise ValueError('`skip` must be non-negative.')
    if _MAX_POSITIVE - num_results <= skip:
      raise ValueError('Skip too large. Should be smaller than '
                       f'{_MAX_POSITIVE} - num_results')
  return _fill(dim, num_results, skip, 1, dtype or np.float32)


def sample_integers(dim, num_results, skip=0):
  """int32 integer points (before the division by 2^num_digits)."""
  return _fill(int(dim), int(num_results), int(skip), 0, np.float32)


def sample_normal(dim, num_results, skip=0, dtype=None):
  """sqrt(2) erfinv(2 u - 1) of the Sobol points (multivariate_normal.py:420)."""
  return _fill(int(dim), int(num_results), int(skip), 2, dtype or np.float32)
