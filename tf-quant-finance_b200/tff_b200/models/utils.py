"""Time grids and draw tensors (`tf_quant_finance/models/utils.py`).

`prepare_grid` is host logic (numpy): it decides the number of Euler steps,
hence the Sobol dimension count and the Philox stream length, and is restated
with TensorFlow's `tf.range` / `tf.linspace` / `tf.searchsorted` semantics.
`generate_mc_normal_draws` produces the `[steps] + batch + [paths, draws]`
tensor on the device for callers that want it; the fused path kernels never
materialise it.
"""
import numpy as np

from tff_b200 import _tensor
from tff_b200.math import random


def _tf_range(start, limit, delta, dtype):
  """`tf.range` for floats: ceil(|limit-start|/|delta|) points, filled by
  repeated addition like TF's CPU kernel (tensorflow==2.12 sequence_ops.cc)."""
  t = np.dtype(dtype).type
  start, limit, delta = t(start), t(limit), t(delta)
  size = int(np.ceil(np.abs((limit - start) / delta)))
  out = np.empty(max(size, 0), dtype=dtype)
  val = start
  for i in range(size):
    out[i] = val
    val = t(val + delta)
  return out


def _tf_linspace(start, stop, num, dtype):
  """`tf.linspace`: concat(start, start + delta*[1..n-2], stop)[:num]."""
  t = np.dtype(dtype).type
  start, stop, num = t(start), t(stop), int(num)
  n_steps = max(num - 1, 1)
  delta = t((stop - start) / t(n_steps))
  inner = (start + delta * np.arange(1, n_steps).astype(dtype)).astype(dtype)
  return np.concatenate([[start], inner, [stop]]).astype(dtype)[:max(num, 0)]


def _grid_from_time_step(*, times, time_step, dtype, tolerance):
  """`models/utils.py:285-306`."""
  grid = _tf_range(0.0, times[-1], time_step, dtype)
  all_times = np.sort(np.concatenate([times, grid]), kind='stable')
  dt = np.concatenate([np.ones(1, dtype=dtype), all_times[1:] - all_times[:-1]])
  all_times = all_times[dt > tolerance]
  time_indices = np.searchsorted(all_times, times, side='left')
  time_indices = np.minimum(time_indices, all_times.shape[0] - 1)
  # Move the indices left if the requested times were removed as duplicates.
  time_indices = np.where(all_times[time_indices] - times > tolerance,
                          time_indices - 1, time_indices)
  return all_times, time_indices.astype(np.int32)


def _grid_from_num_times(*, times, time_step, num_time_steps, dtype):
  """`models/utils.py:309-320`."""
  t = np.dtype(dtype).type
  uniform_grid = _tf_linspace(t(time_step), times[-1] - t(time_step),
                              max(int(num_time_steps) - times.shape[0], 0),
                              dtype)
  grid = np.sort(np.concatenate([uniform_grid, times]), kind='stable')
  all_times = np.concatenate([np.zeros(1, dtype=dtype), grid]).astype(dtype)
  time_indices = np.searchsorted(all_times, times, side='left')
  return all_times, time_indices.astype(np.int32)


def prepare_grid(*, times, time_step, dtype, tolerance=None,
                 num_time_steps=None, times_grid=None):
  """Prepares the grid of times for path generation (`models/utils.py:209-282`).

  Returns numpy `(all_times, mask, time_indices)`.
  """
  dtype = _tensor.np_dtype(dtype)
  if tolerance is None:
    tolerance = 1e-10 if dtype == np.float64 else 1e-6
  tolerance = dtype.type(tolerance)
  times = _tensor.to_numpy(times, dtype)
  if times_grid is None:
    if num_time_steps is None:
      all_times, time_indices = _grid_from_time_step(
          times=times, time_step=time_step, dtype=dtype, tolerance=tolerance)
    else:
      all_times, time_indices = _grid_from_num_times(
          times=times, time_step=time_step, num_time_steps=num_time_steps,
          dtype=dtype)
  else:
    all_times = _tensor.to_numpy(times_grid, dtype)
    idx = np.searchsorted(all_times, times, side='left')
    idx = np.minimum(idx, all_times.shape[0] - 1)
    # Adjust indices to bring `times` closer to `times_grid`.
    diff_1 = all_times[idx] - times
    diff_2 = all_times[np.maximum(idx - 1, 0)] - times
    time_indices = np.where(np.abs(diff_2) > np.abs(diff_1), idx,
                            np.maximum(idx - 1, 0)).astype(np.int32)
  mask = np.zeros(all_times.shape[0], dtype=np.int64)
  np.add.at(mask, time_indices.astype(np.int64), 1)   # scatter_nd handles dups
  return all_times, mask > 0, time_indices


def generate_mc_normal_draws(num_normal_draws, num_time_steps,
                             num_sample_paths, random_type, batch_shape=None,
                             skip=0, seed=None, dtype=None, name=None):
  """Normal draws of shape `[num_time_steps] + batch_shape + [num_sample_paths,
  num_normal_draws]` on the device (`models/utils.py:20-128`)."""
  del name
  if skip is None:
    skip = 0
  dtype = _tensor.np_dtype(dtype, np.float32)
  batch_shape = tuple(int(b) for b in (batch_shape or ()))
  num_normal_draws = int(num_normal_draws)
  num_time_steps = int(num_time_steps)
  num_sample_paths = int(num_sample_paths)
  total_dimension = np.zeros([num_time_steps * num_normal_draws], dtype=dtype)
  rt = random.RandomType(random_type.value)
  if rt in (random.RandomType.PSEUDO_ANTITHETIC,
            random.RandomType.STATELESS_ANTITHETIC):
    sample_shape = (num_sample_paths,) + batch_shape
    is_antithetic = True
  else:
    sample_shape = batch_shape + (num_sample_paths,)
    is_antithetic = False
  draws = random.mv_normal_sample(sample_shape, mean=total_dimension,
                                  random_type=rt, seed=seed, skip=skip)
  draws = draws.reshape(sample_shape + (num_time_steps, num_normal_draws))
  rank = draws.dim()
  if is_antithetic and rank > 3:
    perm = [rank - 2] + list(range(1, rank - 2)) + [0, rank - 1]
  else:
    perm = [rank - 2] + list(range(rank - 2)) + [rank - 1]
  return draws.permute(perm)
