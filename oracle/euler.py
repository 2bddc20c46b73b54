"""Oracle (test infrastructure): the Euler-Maruyama sampler.

Restates `models/euler_sampling.py`:
  * `sample`      27-332  -> `sample`
  * `_sample`     335-402 (draws are always precomputed, as the reference does
                  for STATELESS*/SOBOL and, by default, for PSEUDO*)
  * `_while_loop` 405-464, `_euler_step` 513-537 -> the loop below.
`drift_fn(t, x)` / `volatility_fn(t, x)` are numpy callables with the
reference's conventions: x is `batch_shape + [num_samples, dim]`, drift has the
same shape, volatility is `... + [dim, dim]`.
"""
import numpy as np

from oracle import draws as draws_lib
from oracle import grid as grid_lib


def sample(dim, drift_fn, volatility_fn, times, time_step=None,
           num_time_steps=None, num_samples=1, initial_state=None,
           random_type=None, seed=None, skip=0, times_grid=None,
           normal_draws=None, tolerance=None, dtype=None, return_grid=False,
           path_range=None, return_extrema=False):
  """`euler_sampling.sample` -> batch_shape + [num_samples, k, dim].

  Oracle extensions: `path_range` (a slice of the paths, see
  `draws.generate_mc_normal_draws`); `return_extrema` also returns the running
  maximum and minimum of state component 0 over the initial state and every
  executed step (what a barrier monitored on all grid points sees)."""
  if dtype is None:
    dtype = np.asarray(times).dtype
    if dtype.kind != 'f':
      dtype = np.float32
  dtype = np.dtype(dtype)
  times = np.asarray(times, dtype=dtype)
  if initial_state is None:
    initial_state = np.zeros(dim, dtype=dtype)
  initial_state = np.asarray(initial_state, dtype=dtype)
  batch_shape = initial_state.shape[:-2]
  k = times.shape[0]
  all_times, keep_mask, time_indices = grid_lib.euler_grid(
      times, dtype=dtype, time_step=time_step, num_time_steps=num_time_steps,
      times_grid=times_grid, tolerance=tolerance)

  if normal_draws is not None:
    normal_draws = np.asarray(normal_draws, dtype=dtype)
    r = normal_draws.ndim
    normal_draws = np.transpose(
        normal_draws, [r - 2] + list(range(r - 2)) + [r - 1])
    num_samples = normal_draws.shape[-2]
    if dim != normal_draws.shape[-1]:
      raise ValueError(
          '`dim` should be equal to `normal_draws.shape[2]` but are '
          '{0} and {1} respectively'.format(dim, normal_draws.shape[-1]))

  dt = all_times[1:] - all_times[:-1]                       # :354
  sqrt_dt = np.sqrt(dt)                                      # :355
  state = initial_state + np.zeros([num_samples, dim], dtype=dtype)   # :357
  steps_num = dt.shape[-1]
  if normal_draws is None:
    normal_draws = draws_lib.generate_mc_normal_draws(
        num_normal_draws=dim, num_time_steps=steps_num,
        num_sample_paths=num_samples, batch_shape=batch_shape,
        random_type=(draws_lib.RandomType.PSEUDO if random_type is None
                     else random_type),
        dtype=dtype, seed=seed, skip=skip, path_range=path_range)
    if path_range is not None:         # oracle extension: a slice of the paths
      num_samples = normal_draws.shape[-2]
      state = initial_state + np.zeros([num_samples, dim], dtype=dtype)

  record = k != 1
  written = 0
  slots = [None] * k
  if record:
    slots[0] = state
  written += int(keep_mask[0])
  i = 0
  xmax = np.array(state[..., 0])
  xmin = np.array(state[..., 0])
  while i < steps_num and written < k:                       # cond_fn :426-431
    t = all_times[i + 1]
    dw = normal_draws[i] * sqrt_dt[i]
    dt_inc = dt[i] * drift_fn(t, state)
    vol = volatility_fn(t, state)
    dw_inc = np.einsum('...ij,...j->...i', vol, dw).astype(dtype)
    state = (state + dt_inc + dw_inc).astype(dtype)
    if return_extrema:
      xmax = np.maximum(xmax, state[..., 0])
      xmin = np.minimum(xmin, state[..., 0])
    if record:
      slots[written] = state
    written += int(keep_mask[i + 1])
    i += 1
  if not record:
    out = np.expand_dims(state, axis=-2)
  else:
    # duplicate request times leave trailing TensorArray slots unwritten: TensorFlow's
    # stack() returns zeros there (element shape known)
    slots = [np.zeros_like(state) if s is None else s for s in slots]
    res = np.stack(slots, axis=0)                 # [k] + batch + [N, dim]
    n = res.ndim
    out = np.transpose(res, list(range(1, n - 1)) + [0, n - 1])
  if return_extrema:
    return out, xmax, xmin
  if return_grid:
    return out, (all_times, keep_mask, time_indices)
  return out
