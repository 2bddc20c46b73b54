"""Sobol points as a digital net (`math/qmc/sobol.py`): `sobol_sample` (32-129)
and `sobol_generating_matrices` (132-218).  Unlike `tff.math.random.sobol`
the sequence starts at the origin (index 0) and supports digital shifts,
matrix scrambling and the tent transform."""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200.math.qmc import digital_net
from tff_b200.math.qmc import utils
from tff_b200.math.random import sobol as _sobol_data

InvalidArgumentError = ValueError


def sobol_generating_matrices(dim, num_results, num_digits, validate_args=False, dtype=None,
                              name=None):
  """Host integer table `[dim, ceil(log2 num_results)]` (sobol.py:132-218)."""
  del name
  dtype = digital_net._int_dtype(dtype)  # pylint: disable=protected-access
  dim, num_results, num_digits = int(dim), int(num_results), int(num_digits)
  if validate_args:
    if dim <= 0:
      raise InvalidArgumentError('dim must be positive')
    if num_results <= 0:
      raise InvalidArgumentError('num_results must be positive')
    if num_digits <= 0:
      raise InvalidArgumentError('num_digits must be positive')
  log_num_results = utils.ceil_log2_float32(num_results)
  if validate_args and log_num_results >= 32:
    raise InvalidArgumentError('log2(num_results) must be less than 32')
  a, s, m = _sobol_data._joe_kuo()  # pylint: disable=protected-access
  out = np.empty((dim, log_num_results), dtype=np.int64)
  _lib.check(_lib.lib().tqf_qmc_sobol_generating_matrices(
      a.ctypes.data, s.ctypes.data, m.ctypes.data, a.shape[0], dim, log_num_results, num_digits,
      out.ctypes.data))
  return out.astype(dtype)


def sobol_sample(dim, num_results, sequence_indices=None, digital_shift=None,
                 scrambling_matrices=None, apply_tent_transform=False, validate_args=False,
                 dtype=None, name=None):
  """`[num_results, dim]` Sobol points starting at index 0 (sobol.py:32-129)."""
  del name
  dtype = _tensor.np_dtype(dtype, np.float32)
  num_digits = utils.ceil_log2_float32(num_results)
  g = sobol_generating_matrices(dim, num_results, num_digits, validate_args=validate_args,
                                dtype=np.int32)
  if scrambling_matrices is not None:
    g = digital_net.scramble_generating_matrices(g, scrambling_matrices, num_digits,
                                                 validate_args=validate_args)
  return digital_net.digital_net_sample(
      g, num_results, num_digits, sequence_indices=sequence_indices, digital_shift=digital_shift,
      apply_tent_transform=apply_tent_transform, validate_args=validate_args, dtype=dtype)
