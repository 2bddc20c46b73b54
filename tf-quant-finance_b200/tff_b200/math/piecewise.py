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
    # dtype=None follows `tf.convert_to_tensor(jump_locations)`: arrays / tensors keep
    # their floating type, Python numbers become float32 (`piecewise.py:109-112`)
    self._dtype = _tensor.infer_dtype(jump_locations, dtype)
    self._jump_locations = _tensor.to_numpy(jump_locations, self._dtype)
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
    """Integral over [x1, x2], x1 <= x2 (`piecewise.py:178-208`).

    Batched functions take `x1`, `x2` broadcastable to `batch_shape + [num_points]`.
    """
    del name
    x1 = _tensor.to_numpy(x1, self._dtype)
    x2 = _tensor.to_numpy(x2, self._dtype)
    if self._batch_rank == 0:
      return _integrate(self._jump_locations, self._values, x1, x2)
    batch_shape = self._jump_locations.shape[:-1]
    x1, x2 = np.broadcast_arrays(x1, x2)
    x1 = np.broadcast_to(x1, batch_shape + x1.shape[-1:])
    x2 = np.broadcast_to(x2, batch_shape + x2.shape[-1:])
    out = np.empty(x1.shape + self._values.shape[self._batch_rank + 1:], dtype=self._dtype)
    for b in np.ndindex(*batch_shape):
      out[b] = _integrate(self._jump_locations[b], self._values[b], x1[b], x2[b])
    return out


def _integrate(jump_locations, values, x1, x2):
  """Sum over the pieces of (overlap of [x1, x2] with the piece) * value."""
  lo = np.concatenate([[-np.inf], jump_locations])
  hi = np.concatenate([jump_locations, [np.inf]])
  out = np.zeros(np.broadcast(x1, x2).shape + values.shape[1:], dtype=values.dtype)
  for i in range(values.shape[0]):
    w = np.maximum(np.minimum(x2, hi[i]) - np.maximum(x1, lo[i]), 0)
    out = out + w.reshape(w.shape + (1,) * (values.ndim - 1)) * values[i]
  # (the piece boundaries are float64: a float32 function accumulates in float64 and is rounded once)
  return out.astype(values.dtype)


def find_interval_index(query_xs, interval_lower_xs, last_interval_is_closed=False, dtype=None,
                        name=None):
  """Index of the half-open interval `[x_i, x_{i+1})` each query lies in, -1 before `x_0`
  (`piecewise.py:211-281`); int32 like the reference."""
  del name
  query_xs = _tensor.to_numpy(query_xs, None if dtype is None else _tensor.np_dtype(dtype))
  lower = _tensor.to_numpy(interval_lower_xs, query_xs.dtype)
  indices = np.searchsorted(lower, query_xs, side='right').astype(np.int32) - 1
  if not last_interval_is_closed:
    return indices
  last_index = lower.shape[-1] - 1
  should_cap = (indices == last_index) & (query_xs <= lower[last_index])
  return np.minimum(indices, last_index - should_cap.astype(np.int32)).astype(np.int32)


def convert_to_tensor_or_func(x, dtype=None, name=None):
  """`piecewise.py:421`: (value or function, is_constant)."""
  del name
  if isinstance(x, PiecewiseConstantFunc):
    return x, False
  return _tensor.to_numpy(x, None if dtype is None else _tensor.np_dtype(dtype)), True
