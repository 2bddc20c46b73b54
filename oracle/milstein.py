"""Oracle (test infrastructure): the one-dimensional Milstein sampler.

Restates `models/milstein_sampling.py`:
  * `sample`         35-256  (argument handling, `utils.prepare_grid`)
  * `_sample`        258-353 (precomputed draws: `dim + 3 * dim * stratonovich_order`
                     normals per step, the first `dim` of them drive the path)
  * `_while_loop` / `_milstein_step` 356-425, 598-672 (coefficients at times[i + 1])
  * `_milstein_1d`   565-575.
`drift_fn(t, x)` / `volatility_fn(t, x)` are numpy callables with the reference's
conventions; `grad_volatility_fn(t, x)` returns dS/dx with the shape of the
volatility ([N, 1, 1]).  dim > 1 (Stratonovich integrals) is not restated.
"""
import numpy as np

from oracle import draws as draws_lib
from oracle import grid as grid_lib


def sample(*, dim, drift_fn, volatility_fn, grad_volatility_fn, times, time_step=None,
           num_time_steps=None, num_samples=1, initial_state=None, random_type=None,
           seed=None, skip=0, stratonovich_order=5, dtype=None):
  if dim != 1:
    raise NotImplementedError('the oracle restates the 1-d Milstein scheme only')
  dtype = np.dtype(np.asarray(times).dtype if dtype is None else dtype)
  times = np.asarray(times, dtype=dtype)
  k = times.shape[0]
  if num_time_steps is not None and time_step is not None:
    raise ValueError('Only one of either `num_time_steps` or `time_step` '
                     'should be defined but not both')
  if time_step is None:
    if num_time_steps is None:
      raise ValueError('Either `num_time_steps` or `time_step` should be defined.')
    time_step = dtype.type(times[-1] / dtype.type(num_time_steps))
  all_times, keep_mask, _ = grid_lib.prepare_grid(
      times=times, time_step=dtype.type(time_step), num_time_steps=num_time_steps, dtype=dtype)
  if initial_state is None:
    initial_state = np.zeros(dim, dtype=dtype)
  dt = all_times[1:] - all_times[:-1]
  sqrt_dt = np.sqrt(dt)
  state = np.asarray(initial_state, dtype=dtype) + np.zeros([num_samples, dim], dtype=dtype)
  steps_num = dt.shape[-1]
  all_draws = draws_lib.generate_mc_normal_draws(
      num_normal_draws=dim + 3 * dim * stratonovich_order, num_time_steps=steps_num,
      num_sample_paths=num_samples,
      random_type=draws_lib.RandomType.PSEUDO if random_type is None else random_type,
      dtype=dtype, seed=seed, skip=skip)
  normal_draws = all_draws[:, :, :dim]
  record = k != 1
  slots = [None] * k
  written = 0
  if record:
    slots[0] = state
  written += int(keep_mask[0])
  i = 0
  while i < steps_num and written < k:
    t = all_times[i + 1]
    dw = normal_draws[i] * sqrt_dt[i]
    drift = drift_fn(t, state)
    vol = volatility_fn(t, state)
    grad_vol = grad_volatility_fn(t, state)
    dt_inc = dt[i] * drift
    dw_inc = np.einsum('...ij,...j->...i', np.broadcast_to(vol, state.shape + (1,)), dw)
    hot_vol = np.squeeze(np.broadcast_to(vol * grad_vol, state.shape + (1,)), -1)
    hot_dw = dw * dw - dt[i]
    hot_inc = hot_vol * hot_dw / 2
    state = (state + dt_inc + dw_inc + hot_inc).astype(dtype)
    if record:
      slots[written] = state
    written += int(keep_mask[i + 1])
    i += 1
  if not record:
    return np.expand_dims(state, axis=-2)
  return np.transpose(np.stack(slots, axis=0), [1, 0, 2])
