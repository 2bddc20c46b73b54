"""TensorFlow-compatible Philox normal streams generated on the device.

Replaces `tf.random.stateless_normal(..., alg='philox')` and the first call of
`tf.random.normal(..., seed=)` used by
`math/random_ops/multivariate_normal.py:260-269`.
"""
import ctypes as C
import secrets

import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor


def stateless_key_counter(seed):
  """(key[2], counter[4]) uint32 arrays for an integer seed pair."""
  seed = _tensor.to_numpy(seed).astype(np.int64).reshape(-1)
  if seed.shape[0] != 2:
    raise ValueError('`seed` must be an integer Tensor of shape [2]')
  s = (C.c_int64 * 2)(int(seed[0]), int(seed[1]))
  key = (C.c_uint32 * 2)()
  ctr = (C.c_uint32 * 4)()
  _lib.check(_lib.lib().tqf_philox_stateless_key_counter(s, key, ctr))
  return key, ctr


def stateful_key_counter(seed):
  """Key / counter of a fresh `tf.random.normal(seed=seed)` kernel.

  `seed=None` draws a fresh seed from the OS (TF: non-deterministic)."""
  if seed is None:
    seed = secrets.randbits(31)
  key = (C.c_uint32 * 2)()
  ctr = (C.c_uint32 * 4)()
  _lib.check(_lib.lib().tqf_philox_stateful_key_counter(int(seed), key, ctr))
  return key, ctr


def _fill(key, ctr, shape, dtype, first_element=0, uniform=False):
  dtype = _tensor.np_dtype(dtype)
  shape = tuple(int(s) for s in np.asarray(shape).reshape(-1))
  n = int(np.prod(shape)) if shape else 1
  out = _tensor.empty((n,), dtype)
  fill = _lib.lib().tqf_philox_uniform_fill if uniform else _lib.lib().tqf_philox_normal_fill
  _lib.check(fill(
      key, ctr, first_element, n, _tensor.tqf_dtype(dtype), out.data_ptr(),
      _tensor.current_stream_ptr()))
  return out.reshape(shape)


def stateless_normal(shape, seed, dtype=np.float32):
  """`tf.random.stateless_normal(shape, seed=seed, dtype=dtype, alg='philox')`."""
  key, ctr = stateless_key_counter(seed)
  return _fill(key, ctr, shape, dtype)


def normal(shape, dtype=np.float32, seed=None):
  """First invocation of `tf.random.normal(shape, dtype=dtype, seed=seed)`."""
  key, ctr = stateful_key_counter(seed)
  return _fill(key, ctr, shape, dtype)


def stateless_uniform(shape, seed, dtype=np.float32):
  """`tf.random.stateless_uniform(shape, seed=seed, dtype=dtype, alg='philox')` on [0, 1)."""
  key, ctr = stateless_key_counter(seed)
  return _fill(key, ctr, shape, dtype, uniform=True)


def uniform(shape, dtype=np.float32, seed=None):
  """First invocation of `tf.random.uniform(shape, dtype=dtype, seed=seed)` on [0, 1)."""
  key, ctr = stateful_key_counter(seed)
  return _fill(key, ctr, shape, dtype, uniform=True)


def raw_words(key, counter, first_group, num_groups):
  """uint32 [num_groups, 4] Philox4x32-10 output (bit-exactness tests)."""
  import torch  # pylint: disable=g-import-not-at-top
  out = torch.empty((int(num_groups), 4), dtype=torch.int32,
                    device=_tensor.device())
  k = (C.c_uint32 * 2)(*[int(x) for x in key])
  c = (C.c_uint32 * 4)(*[int(x) for x in counter])
  _lib.check(_lib.lib().tqf_philox_raw_fill(
      k, c, int(first_group), int(num_groups), out.data_ptr(),
      _tensor.current_stream_ptr()))
  return out.view(torch.uint32)
