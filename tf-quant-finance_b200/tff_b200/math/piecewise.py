"""Host-side piecewise constant functions (model parameters).

Mirrors the part of `tf_quant_finance/math/piecewise.py:19-208` the samplers
use: parameters are evaluated once per grid point ON THE HOST and shipped to
the device as per-step coefficient tables, so nothing here runs on the GPU.
"""
import numpy as np

from tff_b200 import _tensor


class PiecewiseConstantFunc:
  """Left-continuous piecewise constant function (`piecewise.py:19-176`).

  f(x) = values[..., i] for jump_locations[..., i-1] < x <= jump_locations[..., i].
  """

  def __init__(self, jump_locations, values, dtype=None, name=None):
    self._name = name or 'PiecewiseConstantFunc'
    self.is_piecewise_constant = True
    self._jump_locations = _tensor.to_numpy(
        jump_locations, None if dtype is None else _tensor.np_dtype(dtype))
    if self._jump_locations.dtype.kind != 'f':
      self._jump_locations = self._jump_locations.astype(np.float32)
    self._dtype = self._jump_locations.dtype
    self._values = _tensor.to_numpy(values, self._dtype)
    self._batch_rank = self._jump_locations.ndim - 1
    sv, sj = list(self._values.shape), list(self._jump_locations.shape)
    if sv[:self._batch_rank] != sj[:-1]:
      raise ValueError(
          'Batch shapes of `values` and `jump_locations` should '
          'be the same but are {0} and {1}'.format(sv[:-1], sj[:-1]))
    if sv[self._batch_rank] - 1 != sj[-1]:
      raise ValueError('Event shape of `values` should have one more '
                       'element than the event shape of `jump_locations` '
                       'but are {0} and {1}'.format(sv[-1], sj[-1]))

  def dtype(self):
    return self._dtype

  def values(self):
    return self._values

  def jump_locations(self):
    return self._jump_locations

  def name(self):
    return self._name

  def __call__(self, x, left_continuous=True, name=None):
    del name
    x = _tensor.to_numpy(x, self._dtype)
    side = 'left' if left_continuous else 'right'
    if self._batch_rank == 0:
      idx = np.searchsorted(self._jump_locations, x, side=side)
      return self._values[idx]
    # batched: x broadcast to batch_shape + [num_points]
    batch_shape = self._jump_locations.shape[:-1]
    x = np.broadcast_to(x, batch_shape + x.shape[-1:])
    out = np.empty(x.shape + self._values.shape[self._batch_rank + 1:],
                   dtype=self._dtype)
    for b in np.ndindex(*batch_shape):
      idx = np.searchsorted(self._jump_locations[b], x[b], side=side)
      out[b] = self._values[b][idx]
    return out

  def integrate(self, x1, x2, name=None):
    """Integral over [x1, x2], x1 <= x2 (`piecewise.py:178-208`), batch-free."""
    del name
    if self._batch_rank != 0:
      raise NotImplementedError('batched integrate is not needed on this path')
    x1 = _tensor.to_numpy(x1, self._dtype)
    x2 = _tensor.to_numpy(x2, self._dtype)
    lo = np.concatenate([[-np.inf], self._jump_locations])
    hi = np.concatenate([self._jump_locations, [np.inf]])
    out = np.zeros(np.broadcast(x1, x2).shape + self._values.shape[1:],
                   dtype=self._dtype)
    for i in range(self._values.shape[0]):
      w = np.maximum(np.minimum(x2, hi[i]) - np.maximum(x1, lo[i]), 0)
      out = out + w.reshape(w.shape + (1,) * (self._values.ndim - 1)) * self._values[i]
    return out


def convert_to_tensor_or_func(x, dtype=None, name=None):
  """`piecewise.py:421`: (value or function, is_constant)."""
  del name
  if isinstance(x, PiecewiseConstantFunc):
    return x, False
  return _tensor.to_numpy(x, None if dtype is None else _tensor.np_dtype(dtype)), True
