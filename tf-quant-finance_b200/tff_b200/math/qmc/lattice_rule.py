"""Rank-1 lattice rules (`math/qmc/lattice_rule.py`): `random_scrambling_vectors`
(40-96) and `lattice_rule_sample` (99-229), sampled by
`tqf_qmc_lattice_rule_fill`."""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200.math.qmc import digital_net
from tff_b200.math.random import philox

InvalidArgumentError = ValueError


def random_scrambling_vectors(dim, seed, validate_args=False, dtype=None, name=None):
  """`[dim]` stateless uniforms on [0, 1) to be used as `additive_shift`."""
  del name
  if validate_args and int(dim) <= 0:
    raise InvalidArgumentError('dim must be positive')
  return philox.stateless_uniform((int(dim),), seed, dtype=_tensor.np_dtype(dtype, np.float32))


def lattice_rule_sample(generating_vectors, dim, num_results, sequence_indices=None,
                        additive_shift=None, apply_tent_transform=False, validate_args=False,
                        dtype=None, name=None):
  """`[num_results, dim]` points `frac(i z / n + shift)` (lattice_rule.py:99-229)."""
  del name
  gv = _tensor.to_numpy(generating_vectors)
  if gv.dtype.kind not in 'iu':
    raise ValueError('generating_vectors must be an integer tensor')
  int_dtype = digital_net._int_dtype(gv.dtype)  # pylint: disable=protected-access
  real_dtype = _tensor.np_dtype(dtype, np.float32)
  dim, num_results = int(dim), int(num_results)
  if validate_args:
    if gv.ndim != 1:
      raise InvalidArgumentError('generating_vectors must have rank 1')
    if dim > gv.size:
      raise InvalidArgumentError('dim must not exceed the size of generating_vectors')
    if num_results <= 0:
      raise InvalidArgumentError('num_results must be positive')
  if gv.ndim != 1 or dim > gv.size:
    raise ValueError('generating_vectors must be a vector with at least `dim` entries')
  gv64 = np.ascontiguousarray(gv[:dim], dtype=np.int64)
  shift = None
  if additive_shift is not None:
    shift = _tensor.to_numpy(additive_shift).astype(real_dtype).reshape(-1)[:dim]
    if shift.size != dim:
      raise ValueError('additive_shift must have at least `dim` entries')
    shift = np.ascontiguousarray(shift, dtype=np.float64)
  seq, count = digital_net._sequence_indices(sequence_indices, False, num_results)  # pylint: disable=protected-access
  out = _tensor.empty((count, dim), real_dtype)
  _lib.check(_lib.lib().tqf_qmc_lattice_rule_fill(
      gv64.ctypes.data, dim, num_results, None if shift is None else shift.ctypes.data,
      None if seq is None else seq.data_ptr(), 0, count, 8 * int_dtype.itemsize,
      int(bool(apply_tent_transform)), _tensor.tqf_dtype(real_dtype), out.data_ptr(),
      _tensor.current_stream_ptr()))
  return out
