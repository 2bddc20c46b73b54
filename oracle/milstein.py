"""Oracle (test infrastructure): the Milstein sampler, one- and multi-dimensional.

Restates `models/milstein_sampling.py`:
  * `sample`         35-256  (argument handling, `utils.prepare_grid`)
  * `_sample`        258-353 (precomputed draws: `dim + 3 * dim * stratonovich_order`
                     normals per step, the first `dim` of them drive the path)
  * `_while_loop` / `_milstein_step` 356-425, 598-672 (coefficients at times[i + 1])
  * `_milstein_1d`   565-575;
  * dim > 1: `_stratonovich_integral` 481-527, `_milstein_hot` 530-551,
    `_stratonovich_drift_update` 554-562, `_milstein_nd` 578-595, the auxiliary
    draws of `_milstein_step` 613-618 -- restated as the code is written (index
    placement of `lbar` and the missing 1/2 of the drift update included).
`drift_fn(t, x)` / `volatility_fn(t, x)` are numpy callables with the reference's
conventions.  dim 1: `grad_volatility_fn(t, x)` returns dS/dx with the shape of
the volatility ([N, 1, 1]).  dim > 1: `grad_volatility_fn(t, x)` returns the list
over l of dS/dx_l, each [N, dim, dim] (what the reference builds from the
forward gradients along the unit vectors, 229-235, 642-650).
"""
import numpy as np

from oracle import draws as draws_lib
from oracle import grid as grid_lib


def sample(*, dim, drift_fn, volatility_fn, grad_volatility_fn, times, time_step=None,
           num_time_steps=None, num_samples=1, initial_state=None, random_type=None,
           seed=None, skip=0, stratonovich_order=5, dtype=None):
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
  aux = []
  start = dim
  for _ in range(3):                                   # milstein_sampling.py:296-300
    aux.append(all_draws[:, :, start:start + dim * stratonovich_order])
    start += dim * stratonovich_order
  record = k != 1
  slots = [None] * k
  written = 0
  if record:
    slots[0] = state
  written += int(keep_mask[0])
  i = 0
  while i < steps_num and written < k:
    t = all_times[i + 1]
    if dim > 1:
      strat = [a[i].reshape(num_samples, dim, stratonovich_order) for a in aux]
      state = _milstein_nd(dim, num_samples, normal_draws[i], dt[i], sqrt_dt[i], state,
                           drift_fn(t, state), volatility_fn(t, state),
                           grad_volatility_fn(t, state), strat, stratonovich_order).astype(dtype)
      if record:
        slots[written] = state
      written += int(keep_mask[i + 1])
      i += 1
      continue
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


def _outer(v1, v2):
  return np.einsum('...i,...j->...ij', v1, v2)


def _stratonovich_integral(dim, dt, sqrt_dt, dw, draws, order):
  """milstein_sampling.py:481-527: approximate J(i, j), [N, dim, dim]."""
  p = order - 1
  sqrt_rho_p = np.sqrt(np.asarray(
      1 / 12 - sum(1 / r**2 for r in range(1, order + 1)) / 2 / np.pi**2, dtype=dw.dtype))
  mu = draws[0]
  zeta = np.transpose(draws[1], [2, 0, 1])             # [order, N, dim]
  eta = np.transpose(draws[2], [2, 0, 1])
  xi = dw / sqrt_dt
  r_i = np.stack([np.ones(zeta[0].shape + (dim,), dtype=zeta.dtype) / r
                  for r in range(1, order + 1)], 0)
  value = dt * (_outer(dw, dw) / 2 + sqrt_rho_p * (_outer(mu[..., p], xi) - _outer(xi, mu[..., p])))
  y = np.sqrt(np.asarray(2, dtype=dw.dtype)) * xi + eta
  value = value + dt * np.sum((_outer(zeta, y) - _outer(y, zeta)) * r_i, 0) / (2 * np.pi)
  return value


def _milstein_hot(dim, vol, grad_vol, dt, sqrt_dt, dw, draws, order):
  """milstein_sampling.py:530-551."""
  integrals = _stratonovich_integral(dim, dt, sqrt_dt, dw, draws, order)
  idx = np.arange(dim)
  integrals[:, idx, idx] = dw * dw / 2                  # tf.linalg.set_diag
  stacked = []
  for state_ix in range(dim):
    stacked.append(np.transpose(np.stack([g[..., state_ix, :] for g in grad_vol], -1), [0, 2, 1]))
  stacked = np.stack(stacked, 0)                        # [dim, N, dim, dim]
  lbar = np.matmul(stacked, vol)
  return np.transpose(np.sum(lbar * integrals, axis=(-2, -1)))


def _milstein_nd(dim, num_samples, dw, dt, sqrt_dt, state, drift, vol, grad_vol, draws, order):
  """milstein_sampling.py:578-595 (+ 554-562)."""
  vol = np.broadcast_to(vol, (num_samples, dim, dim)).astype(state.dtype)
  grad_vol = [np.broadcast_to(g, (num_samples, dim, dim)).astype(state.dtype) for g in grad_vol]
  dw = dw * sqrt_dt
  drift_update = np.einsum('nkm,nm->nk', np.concatenate(grad_vol, 2), vol.reshape(num_samples, -1))
  dt_inc = dt * (drift - drift_update)
  dw_inc = np.einsum('nij,nj->ni', vol, dw)
  hot_inc = _milstein_hot(dim, vol, grad_vol, dt, sqrt_dt, dw, draws, order)
  return state + dt_inc + dw_inc + hot_inc
