"""`tff.math.random.halton` on the device: the NON-randomized Halton sequence
(`math/random_ops/halton/halton_impl.py:59-288` with `randomized=False`).

The Owen-scrambled variant (`randomized=True`, the reference's default) draws
its permutations with TensorFlow's random shuffle and is not implemented: it
raises `NotImplementedError`.
"""
import ctypes as C

import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor

_MAX_DIMENSION = 1000
_MAX_INDEX_BY_DTYPE = {np.dtype(np.float32): 2**24 - 1, np.dtype(np.float64): 2**53 - 1}


def _first_primes(n):
  sieve = np.ones(8000, dtype=bool)
  sieve[:2] = False
  for i in range(2, 90):
    if sieve[i]:
      sieve[i * i::i] = False
  return np.flatnonzero(sieve)[:n].astype(np.int32)      # the 1000th prime is 7919


def _tables(dim, dtype):
  """radixes [dim], digits per axis [dim], weights [dim, max_size] (halton_impl.py:250-273)."""
  radixes = _first_primes(dim)
  # evaluated in float64 on a Python int, as the reference does at import time (530-534)
  sizes = (np.floor(np.log(_MAX_INDEX_BY_DTYPE[dtype]) / np.log(radixes)) + 1).astype(np.int32)
  max_size = int(sizes.max())
  exponents = np.tile(np.arange(max_size, dtype=dtype)[None, :], [dim, 1])
  capped = np.where(exponents >= sizes[:, None].astype(dtype), np.zeros_like(exponents), exponents)
  weights = np.round(radixes.astype(dtype)[:, None]**capped).astype(dtype)
  return radixes, sizes, np.ascontiguousarray(weights, dtype=np.float64), max_size


def _fill(dim, first_index, count, kind, dtype):
  dtype = _tensor.np_dtype(np.float32 if dtype is None else dtype)
  dim = int(dim)
  if dim < 1 or dim > _MAX_DIMENSION:
    raise ValueError('`dim` should be in [1, {}]'.format(_MAX_DIMENSION))
  if first_index + count > _MAX_INDEX_BY_DTYPE[dtype]:
    raise ValueError('Maximum sequence index exceeded. Maximum index for dtype %s is %d.'
                     % (dtype, _MAX_INDEX_BY_DTYPE[dtype]))
  radixes, sizes, weights, max_size = _tables(dim, dtype)
  out = _tensor.empty((int(count), dim), dtype)
  _lib.require_cuda()
  _lib.check(_lib.lib().tqf_halton_fill(
      weights.ctypes.data, sizes.ctypes.data, radixes.ctypes.data, dim, max_size,
      int(first_index), int(count), kind, _tensor.tqf_dtype(dtype), out.data_ptr(),
      _tensor.current_stream_ptr()))
  return out


def _range_of(num_results, sequence_indices):
  if (num_results is None) == (sequence_indices is None):
    raise ValueError('Either `num_results` or `sequence_indices` must be'
                     ' specified but not both.')
  if sequence_indices is None:
    return 0, int(_tensor.to_numpy(num_results))
  idx = np.asarray(_tensor.to_numpy(sequence_indices)).astype(np.int64).reshape(-1)
  if idx.size and not np.all(np.diff(idx) == 1):
    raise NotImplementedError(
        'the B200 Halton kernel generates contiguous index ranges '
        '(sequence_indices = range(start, start + n)) only')
  return (int(idx[0]) if idx.size else 0), int(idx.size)


def sample(dim, num_results=None, sequence_indices=None, randomized=True,
           randomization_params=None, seed=None, validate_args=False, dtype=None, name=None):
  """`halton.sample`: returns `(samples [n, dim] CUDA tensor, None)` for `randomized=False`."""
  del validate_args, name, seed
  if randomized or randomization_params is not None:
    raise NotImplementedError(
        'The randomized (Owen-scrambled) Halton sequence is not implemented by the B200 '
        'engine; pass randomized=False (SURVEY 8f-4).')
  first, count = _range_of(num_results, sequence_indices)
  return _fill(dim, first, count, 1, dtype), None


def sample_normal(dim, num_results, skip=0, dtype=None):
  """`sqrt(2) erfinv(2 u - 1)` of Halton points `skip .. skip + num_results - 1`
  (`multivariate_normal.py:395-420`), fused in the fill kernel."""
  return _fill(dim, int(skip), int(num_results), 2, dtype)
