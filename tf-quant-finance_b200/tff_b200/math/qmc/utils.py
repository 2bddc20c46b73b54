"""`math/qmc/utils.py:23-158` for the small integer tables of the QMC samplers.

These helpers act on host-side (numpy) tables of at most `[dim, 63]` integers:
the generating / scrambling matrices.  Tensors of points never pass through
them; those are produced on the device (`tqf_qmc_digital_net_fill`,
`tqf_qmc_lattice_rule_fill`).
"""
import numpy as np

from tff_b200 import _tensor


def _int_array(value, dtype=None):
  arr = _tensor.to_numpy(value)
  if dtype is not None:
    return arr.astype(_tensor.np_dtype(dtype))
  if arr.dtype.kind not in 'iu':
    arr = arr.astype(np.int32)
  return arr


def exp2(value):
  """`2 ** value`, saturated at the integer type's maximum (utils.py:23-54)."""
  value = _int_array(value)
  dtype = value.dtype
  limit = 8 * dtype.itemsize - (0 if dtype.kind == 'u' else 1)
  safe = np.where(value >= limit, 0, value).astype(dtype)
  return np.where(value >= limit, np.iinfo(dtype).max,
                  np.left_shift(np.ones_like(value), safe)).astype(dtype)


def log2(value):
  """`log(value) / log(2)` in the dtype of `value` (utils.py:57-77)."""
  value = _tensor.to_numpy(value)
  return np.log(value) / np.log(np.asarray(2, dtype=value.dtype))


def ceil_log2_float32(num_results):
  """`ceil(log2(float32(num_results)))` the way every sampler of this package
  sizes its index bits (digital_net.py:318-320, sobol.py:96-98, 183-185)."""
  return int(np.ceil(log2(np.float32(int(num_results)))))


def get_shape(value):
  """utils.py:80-91."""
  return tuple(value.shape)


def tent_transform(value):
  """`where(value < 0.5, 2 value, 2 (1 - value))` (utils.py:94-117).

  Accepts a device tensor (the samplers apply it in-kernel instead)."""
  import torch  # pylint: disable=g-import-not-at-top
  if isinstance(value, torch.Tensor):
    return torch.where(value < 0.5, 2 * value, 2 * (1 - value))
  value = np.asarray(value)
  return np.where(value < 0.5, 2 * value, 2 * (1 - value)).astype(value.dtype)


def filter_tensor(value, bit_mask, bit_index):
  """`value` where bit `bit_index` of `bit_mask` is set, else 0 (utils.py:120-158)."""
  value = _int_array(value)
  bit_mask = _int_array(bit_mask, value.dtype)
  bit_index = _int_array(bit_index, value.dtype)
  is_set = (np.right_shift(bit_mask, bit_index) & 1) == 1
  return np.where(is_set, value, 0).astype(value.dtype)
