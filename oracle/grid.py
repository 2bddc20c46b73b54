"""Oracle (test infrastructure): time grids of the Euler sampler.

Restates `models/utils.py:209-320` (`prepare_grid`, `_grid_from_time_step`,
`_grid_from_num_times`) together with the TensorFlow ops they call:

* `tf.range(start, limit, delta)` for floats: `size = ceil(|limit-start|/|delta|)`
  and the CPU kernel fills by repeated addition `val += delta`
  (tensorflow==2.12 `core/kernels/sequence_ops.cc`, RangeFunctor<CPUDevice>).
* `tf.linspace(start, stop, num)` (`math_ops.linspace_nd`):
  `concat(start, start + delta * [1 .. n_steps-1], stop)[:num]` with
  `delta = (stop - start) / max(num - 1, 1)`.
* `tf.sort`, `tf.searchsorted(side='left')`, `tf.boolean_mask`, `tf.scatter_nd`.
"""
import numpy as np


def tf_range(start, limit, delta, dtype):
  dtype = np.dtype(dtype).type
  start, limit, delta = dtype(start), dtype(limit), dtype(delta)
  size = int(np.ceil(np.abs((limit - start) / delta)))
  out = np.empty(max(size, 0), dtype=dtype)
  val = start
  for i in range(size):
    out[i] = val
    val = dtype(val + delta)
  return out


def tf_linspace(start, stop, num, dtype):
  dtype = np.dtype(dtype).type
  start, stop = dtype(start), dtype(stop)
  num = int(num)
  n_steps = max(num - 1, 1)
  delta = dtype((stop - start) / dtype(n_steps))
  inner = start + delta * np.arange(1, n_steps, dtype=np.int64).astype(dtype)
  full = np.concatenate([[start], inner.astype(dtype), [stop]]).astype(dtype)
  return full[:max(num, 0)]


def grid_from_time_step(times, time_step, dtype, tolerance):
  """`models/utils.py:285-306`."""
  times = np.asarray(times, dtype=dtype)
  grid = tf_range(0.0, times[-1], time_step, dtype)
  all_times = np.sort(np.concatenate([times, grid]), kind='stable')
  dt = all_times[1:] - all_times[:-1]
  dt = np.concatenate([np.ones(1, dtype=dtype), dt])
  all_times = all_times[dt > tolerance]
  idx = np.searchsorted(all_times, times, side='left').astype(np.int32)
  idx = np.minimum(idx, all_times.shape[0] - 1)
  idx = np.where(all_times[idx] - times > tolerance, idx - 1, idx)
  return all_times, idx.astype(np.int32)


def grid_from_num_times(times, time_step, num_time_steps, dtype):
  """`models/utils.py:309-320`."""
  times = np.asarray(times, dtype=dtype)
  ts = np.dtype(dtype).type(time_step)
  uniform = tf_linspace(ts, times[-1] - ts,
                        max(int(num_time_steps) - times.shape[0], 0), dtype)
  grid = np.sort(np.concatenate([uniform, times]), kind='stable')
  all_times = np.concatenate([np.zeros(1, dtype=dtype), grid]).astype(dtype)
  idx = np.searchsorted(all_times, times, side='left').astype(np.int32)
  return all_times, idx


def prepare_grid(*, times, time_step, dtype, tolerance=None,
                 num_time_steps=None, times_grid=None):
  """`models/utils.py:209-282`.  Returns (all_times, mask, time_indices)."""
  dtype = np.dtype(dtype)
  if tolerance is None:
    tolerance = 1e-10 if dtype == np.float64 else 1e-6
  tolerance = dtype.type(tolerance)
  times = np.asarray(times, dtype=dtype)
  if times_grid is None:
    if num_time_steps is None:
      all_times, idx = grid_from_time_step(times, time_step, dtype, tolerance)
    else:
      all_times, idx = grid_from_num_times(times, time_step, num_time_steps,
                                           dtype)
  else:
    all_times = np.asarray(times_grid, dtype=dtype)
    idx = np.searchsorted(all_times, times, side='left').astype(np.int32)
    # tf.gather on CPU raises for an out-of-range index; the reference is only
    # used with times <= times_grid[-1] here, so clip defensively.
    idx_c = np.minimum(idx, all_times.shape[0] - 1)
    d1 = all_times[idx_c] - times
    d2 = all_times[np.maximum(idx_c - 1, 0)] - times
    idx = np.where(np.abs(d2) > np.abs(d1), idx_c, np.maximum(idx_c - 1, 0))
  mask = np.zeros(all_times.shape[0], dtype=np.int64)
  np.add.at(mask, idx.astype(np.int64), 1)
  return all_times, mask > 0, idx.astype(np.int32)


def euler_grid(times, *, dtype, time_step=None, num_time_steps=None,
               times_grid=None, tolerance=None):
  """The argument handling of `models/euler_sampling.py:232-283`."""
  dtype = np.dtype(dtype)
  times = np.asarray(times, dtype=dtype)
  if tolerance is None:
    tolerance = 1e-10 if dtype == np.float64 else 1e-6
  if num_time_steps is not None and time_step is not None:
    raise ValueError(
        'When `times_grid` is not supplied only one of either '
        '`num_time_steps` or `time_step` should be defined but not both.')
  if times_grid is None:
    if time_step is None:
      if num_time_steps is None:
        raise ValueError(
            'When `times_grid` is not supplied, either `num_time_steps` '
            'or `time_step` should be defined.')
      time_step = dtype.type(times[-1] / dtype.type(num_time_steps))
    else:
      time_step = dtype.type(time_step)
  return prepare_grid(times=times, time_step=time_step,
                      num_time_steps=num_time_steps, times_grid=times_grid,
                      tolerance=tolerance, dtype=dtype)
