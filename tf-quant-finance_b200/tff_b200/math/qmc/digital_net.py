"""Digital nets in base 2 on the device (`math/qmc/digital_net.py`).

`random_digital_shift` (45-95) and `random_scrambling_matrices` (98-152) draw
TensorFlow's stateless integer uniform on the device
(`tqf_philox_uniform_int_fill`); `scramble_generating_matrices` (422-527) is
`[dim, log2 n]` table work done by libtqf on the host;
`digital_net_sample` (205-419) produces the `[num_results, dim]` points with
`tqf_qmc_digital_net_fill`.
"""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200.math.qmc import utils
from tff_b200.math.random import philox

InvalidArgumentError = ValueError


def _int_dtype(dtype):
  dtype = _tensor.np_dtype(dtype, np.int32)
  if dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
    raise ValueError('dtype must be int32 or int64, got {}'.format(dtype))
  return dtype


def _random_stateless_uniform(shape, num_digits, seed, validate_args, dtype):
  """digital_net.py:155-202."""
  dtype = _int_dtype(dtype)
  shape = tuple(int(s) for s in shape)
  num_digits = int(num_digits)
  if validate_args:
    if any(s <= 0 for s in shape):
      raise InvalidArgumentError('shape must be positive')
    if num_digits <= 0:
      raise InvalidArgumentError('num_digits must be positive')
  minval = int(utils.exp2(np.asarray(num_digits - 1, dtype=dtype)))
  maxval = int(utils.exp2(np.asarray(num_digits, dtype=dtype)))
  key, ctr = philox.stateless_key_counter(seed)
  n = int(np.prod(shape))
  out = _tensor.empty((n,), dtype)
  _lib.check(_lib.lib().tqf_philox_uniform_int_fill(
      key, ctr, minval, maxval, n, 8 * dtype.itemsize, out.data_ptr(),
      _tensor.current_stream_ptr()))
  return out.reshape(shape)


def random_digital_shift(dim, num_digits, seed, validate_args=False, dtype=None, name=None):
  """`[dim]` integers in `[2^(num_digits-1), 2^num_digits)` (digital_net.py:45-95)."""
  del name
  return _random_stateless_uniform((int(dim),), num_digits, seed, validate_args, dtype)


def random_scrambling_matrices(dim, num_digits, seed, validate_args=False, dtype=None, name=None):
  """`[dim, num_digits]` such integers (digital_net.py:98-152)."""
  del name
  return _random_stateless_uniform((int(dim), int(num_digits)), num_digits, seed, validate_args,
                                   dtype)


def _host_table(value):
  arr = _tensor.to_numpy(value)
  if arr.dtype.kind not in 'iu':
    raise ValueError('expected an integer tensor, got {}'.format(arr.dtype))
  return arr


def scramble_generating_matrices(generating_matrices, scrambling_matrices, num_digits,
                                 validate_args=False, dtype=None, name=None):
  """Linear matrix scrambling of generating matrices (digital_net.py:422-527).

  Returns a host (numpy) integer table `[dim, num_columns]`."""
  del name
  g = _host_table(generating_matrices)
  s = _host_table(scrambling_matrices)
  dtype = _int_dtype(dtype or g.dtype)
  num_digits = int(num_digits)
  if validate_args:
    if g.ndim != s.ndim:
      raise InvalidArgumentError('input matrices must have the same rank')
    if num_digits <= 0:
      raise InvalidArgumentError('num_digits must be positive')
  if g.ndim != 2 or s.ndim != 2 or g.shape[0] != s.shape[0]:
    raise ValueError('generating_matrices and scrambling_matrices must be [dim, columns] tables')
  g64 = np.ascontiguousarray(g.astype(dtype), dtype=np.int64)
  s64 = np.ascontiguousarray(s.astype(dtype), dtype=np.int64)
  out = np.empty_like(g64)
  _lib.check(_lib.lib().tqf_qmc_scramble_generating_matrices(
      g64.ctypes.data, s64.ctypes.data, g64.shape[0], g64.shape[1], s64.shape[1], num_digits,
      out.ctypes.data))
  return out.astype(dtype)


def _sequence_indices(sequence_indices, validate_args, num_results):
  """Device int64 indices (or None for 0 .. num_results - 1)."""
  import torch  # pylint: disable=g-import-not-at-top
  if sequence_indices is None:
    return None, int(num_results)
  if isinstance(sequence_indices, torch.Tensor):
    seq = sequence_indices
  elif hasattr(sequence_indices, '__dlpack__') and not isinstance(sequence_indices, np.ndarray):
    seq = _tensor.from_dlpack(sequence_indices)
  else:
    seq = torch.as_tensor(np.asarray(_tensor.to_numpy(sequence_indices)))
  if validate_args:
    if seq.dim() != 1:
      raise InvalidArgumentError('sequence_indices must have rank 1')
    if seq.numel() and int(seq.max()) >= int(num_results):
      raise InvalidArgumentError('values in sequence_indices must be less than num_results')
  seq = seq.reshape(-1).to(device=_tensor.device(), dtype=torch.int64).contiguous()
  return seq, int(seq.numel())


def digital_net_sample(generating_matrices, num_results, num_digits, sequence_indices=None,
                       scrambling_matrices=None, digital_shift=None, apply_tent_transform=False,
                       validate_args=False, dtype=None, name=None):
  """`[num_results, dim]` (or `[len(sequence_indices), dim]`) points of the net
  (digital_net.py:205-419)."""
  del name
  g = _host_table(generating_matrices)
  int_dtype = _int_dtype(g.dtype)
  real_dtype = _tensor.np_dtype(dtype, np.float32)
  num_results, num_digits = int(num_results), int(num_digits)
  if validate_args:
    if g.ndim != 2:
      raise InvalidArgumentError('generating_matrices must have rank 2')
    if num_results <= 0:
      raise InvalidArgumentError('num_results must be positive')
    if num_digits <= 0:
      raise InvalidArgumentError('num_digits must be positive')
  if g.ndim != 2:
    raise ValueError('generating_matrices must have rank 2')
  dim = g.shape[0]
  log_num_results = utils.ceil_log2_float32(num_results)
  if validate_args and log_num_results >= 32:
    raise InvalidArgumentError('log2(num_results) must be less than 32')
  shift = None
  if digital_shift is not None:
    shift = _tensor.to_numpy(digital_shift)
    if validate_args:
      if shift.ndim != 1:
        raise InvalidArgumentError('digital_shift must have rank 1')
      if shift.size != dim:
        raise InvalidArgumentError('digital_shift must have size tf.shape(generating_matrices)[0]')
    shift = np.ascontiguousarray(shift.astype(int_dtype).reshape(-1), dtype=np.int64)
    if shift.size != dim:
      raise ValueError('digital_shift must have one entry per coordinate')
  if scrambling_matrices is not None:
    s = _host_table(scrambling_matrices)
    if validate_args and s.shape != g.shape:
      raise InvalidArgumentError('scrambling_matrices must have the same shape as generating_matrices')
    g = scramble_generating_matrices(g, s, num_digits, validate_args=validate_args, dtype=int_dtype)
  seq, count = _sequence_indices(sequence_indices, validate_args, num_results)
  g64 = np.ascontiguousarray(g, dtype=np.int64)
  out = _tensor.empty((count, dim), real_dtype)
  _lib.check(_lib.lib().tqf_qmc_digital_net_fill(
      g64.ctypes.data, dim, g64.shape[1], log_num_results,
      None if shift is None else shift.ctypes.data,
      None if seq is None else seq.data_ptr(), 0, count, num_digits, 8 * int_dtype.itemsize,
      int(bool(apply_tent_transform)), _tensor.tqf_dtype(real_dtype), out.data_ptr(),
      _tensor.current_stream_ptr()))
  return out
