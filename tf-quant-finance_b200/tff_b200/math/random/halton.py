"""`tff.math.random.halton` on the device (`math/random_ops/halton/halton_impl.py:59-322`):
the plain Halton sequence and its Owen-randomized variant (`randomized=True`, the
reference's default).

Randomization follows the reference's construction for an integer `seed`: the digit
permutations are `stateless_random_shuffle(range(p), seed=(seed + i, p))`
(`halton_impl.py:342-379`) -- built by the library on the host from the same Philox
stream TensorFlow's `stateless_uniform` draws -- and the trailing-zero correction is
`stateless_uniform([dim, 1], seed=(seed, seed)) / p^size` (303-322).  With `seed=None`
the reference shuffles with the global generator (not reproducible); here a seed is
drawn from the OS.  The `HaltonParams` returned can be passed back as
`randomization_params`, exactly as in the reference.
"""
import collections
import ctypes as C
import os

import numpy as np
import torch

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


# perms: int32 CUDA tensor [num_coeffs * sum(radixes)]; zero_correction: CUDA tensor [dim]
HaltonParams = collections.namedtuple('HaltonParams', ['perms', 'zero_correction'])
_NUM_COEFFS_BY_DTYPE = {np.dtype(np.float32): 24, np.dtype(np.float64): 54}


def _randomization_params(dim, seed, dtype, randomization_params):
  """`(perms device int32, zero_correction host float64 [dim], HaltonParams)`
  (`halton_impl.py:290-322`)."""
  from tff_b200.math.random import philox  # pylint: disable=g-import-not-at-top
  radixes, sizes, _, _ = _tables(dim, dtype)
  nc = _NUM_COEFFS_BY_DTYPE[dtype]
  radix_sum = int(radixes.astype(np.int64).sum())
  perms, zc = (None, None) if randomization_params is None else randomization_params
  if seed is None and (perms is None or zc is None):
    seed = int.from_bytes(os.urandom(4), 'little') >> 1
  if perms is None:
    host = np.empty(nc * radix_sum, dtype=np.int32)
    _lib.check(_lib.lib().tqf_halton_permutations(
        int(seed), radixes.ctypes.data, int(dim), nc, host.ctypes.data))
    perms = torch.as_tensor(host, device=_tensor.device())
  else:
    perms = _tensor.from_dlpack(perms) if not isinstance(perms, np.ndarray) else torch.as_tensor(perms)
    perms = perms.to(device=_tensor.device(), dtype=torch.int32).reshape(-1).contiguous()
    if int(perms.shape[0]) != nc * radix_sum:
      raise ValueError('randomization_params.perms has {} entries, expected {} for dim={} and '
                       '{}'.format(int(perms.shape[0]), nc * radix_sum, dim, dtype))
  if zc is None:
    u = philox.stateless_uniform([int(dim), 1], (int(seed), int(seed)), dtype=dtype)
    u = u.cpu().numpy().reshape(-1)
    zc_host = (u / (radixes.astype(dtype)**sizes.astype(dtype)).astype(dtype)).astype(dtype)
    zc = torch.as_tensor(zc_host, device=_tensor.device())
  else:
    zc = (torch.as_tensor(zc) if isinstance(zc, np.ndarray) else _tensor.from_dlpack(zc))
    zc = zc.to(device=_tensor.device(), dtype=_tensor.torch_dtype(dtype)).reshape(-1)
    zc_host = zc.cpu().numpy()
  return perms, np.ascontiguousarray(zc_host, dtype=np.float64), HaltonParams(perms, zc)


def _fill(dim, first_index, count, kind, dtype, randomized=False, seed=None,
          randomization_params=None):
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
  if not randomized:
    _lib.check(_lib.lib().tqf_halton_fill(
        weights.ctypes.data, sizes.ctypes.data, radixes.ctypes.data, dim, max_size,
        int(first_index), int(count), kind, _tensor.tqf_dtype(dtype), out.data_ptr(),
        _tensor.current_stream_ptr()))
    return out, None
  perms, zc_host, params = _randomization_params(dim, seed, dtype, randomization_params)
  _lib.check(_lib.lib().tqf_halton_randomized_fill(
      weights.ctypes.data, sizes.ctypes.data, radixes.ctypes.data, dim, max_size,
      perms.data_ptr(), zc_host.ctypes.data, int(first_index), int(count), kind,
      _tensor.tqf_dtype(dtype), out.data_ptr(), _tensor.current_stream_ptr()))
  return out, params


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
  """`halton.sample`: returns `(samples [n, dim] CUDA tensor, HaltonParams or None)`."""
  del validate_args, name
  first, count = _range_of(num_results, sequence_indices)
  return _fill(dim, first, count, 1, dtype, bool(randomized), seed, randomization_params)


def sample_normal(dim, num_results, skip=0, dtype=None, randomized=False, seed=None,
                  randomization_params=None):
  """`sqrt(2) erfinv(2 u - 1)` of Halton points `skip .. skip + num_results - 1`
  (`multivariate_normal.py:391-420`), fused in the fill kernel."""
  return _fill(dim, int(skip), int(num_results), 2, dtype, bool(randomized), seed,
               randomization_params)[0]
