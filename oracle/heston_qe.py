"""Oracle (test infrastructure): Andersen's Quadratic-Exponential scheme.

Restates `models/heston/heston_model.py`: `sample_paths` 177-320,
`_sample_paths` 322-460, `_update_variance` 522-551, `_update_log_spot`
554-572 and `_prepare_grid` 575-639 (duplicates kept, stable argsort).
"""
import numpy as np
from scipy import special

from oracle import draws as draws_lib
from oracle import grid as grid_lib
from oracle import models as models_lib


def _params_at(p, t, dtype):
  if callable(p):
    return np.asarray(p(t), dtype=dtype)
  return np.asarray(p, dtype=dtype) * np.ones_like(t)


def prepare_grid(times, time_step, dtype, params, times_grid=None):
  """`_prepare_grid` (`heston_model.py:575-639`)."""
  times = np.asarray(times, dtype=dtype)
  jumps = [np.asarray(p.jump_locations(), dtype=dtype) for p in params
           if isinstance(p, models_lib.PiecewiseConstantFunc)]
  if times_grid is None:
    grid = grid_lib.tf_range(0.0, times[-1], time_step, dtype)
    all_times = np.concatenate([grid, times] + jumps)
    mask = np.concatenate([np.zeros(grid.shape, bool), np.ones(times.shape, bool)] +
                          [np.zeros(j.shape, bool) for j in jumps])
    perm = np.argsort(all_times, kind='stable')
    return all_times[perm], mask[perm]
  all_times, mask, _ = grid_lib.prepare_grid(times=times, time_step=time_step,
                                             times_grid=times_grid, dtype=dtype)
  return all_times, mask


def _update_variance(kappa, theta, volvol, v, dt, z, psi_c=1.5):
  e = np.exp(-kappa * dt)
  vv2 = volvol**2
  m = theta + (v - theta) * e
  s2 = v * vv2 * e / kappa * (1 - e) + theta * vv2 / 2 / kappa * (1 - e)**2
  psi = s2 / m**2
  u = 0.5 * (1 + special.erf(z / np.sqrt(2.)))
  with np.errstate(all='ignore'):
    psi_inv = 2 / psi
    b2 = psi_inv - 1 + np.sqrt(psi_inv * (psi_inv - 1))
    a = m / (1 + b2)
    v_true = a * (np.sqrt(b2) + z)**2
    p = (psi - 1) / (psi + 1)
    beta = (1 - p) / m
    v_false = np.where(u > p, np.log(1 - p) - np.log(1 - u), 0.0) / beta
  return np.where(psi < psi_c, v_true, v_false)


def _update_log_spot(kappa, theta, volvol, rho, v, v_next, x, dt, z, g1=0.5, g2=0.5):
  k0 = -rho * kappa * theta / volvol * dt
  k1 = g1 * dt * (kappa * rho / volvol - 0.5) - rho / volvol
  k2 = g2 * dt * (kappa * rho / volvol - 0.5) + rho / volvol
  k3 = g1 * dt * (1 - rho**2)
  k4 = g2 * dt * (1 - rho**2)
  return x + k0 + k1 * v + k2 * v_next + np.sqrt(k3 * v + k4 * v_next) * z


def sample_paths(mean_reversion, theta, volvol, rho, times, initial_state,
                 num_samples=1, random_type=None, seed=None, time_step=None,
                 skip=0, tolerance=1e-6, num_time_steps=None, times_grid=None,
                 normal_draws=None, dtype=np.float64, path_range=None,
                 return_extrema=False):
  """`HestonModel.sample_paths` -> [num_samples, k, 2] (log-spot, variance)."""
  dtype = np.dtype(dtype)
  times = np.asarray(times, dtype=dtype)
  x0 = np.asarray(initial_state, dtype=dtype)
  if times_grid is None:
    if time_step is None:
      if num_time_steps is None:
        raise ValueError(
            'When `times_grid` is not supplied, either `num_time_steps` '
            'or `time_step` should be defined.')
      time_step = dtype.type(times[-1] / dtype.type(num_time_steps))
    else:
      if num_time_steps is not None:
        raise ValueError(
            'Both `time_step` and `num_time_steps` can not be `None` '
            'simultaneously when calling sample_paths of HestonModel.')
      time_step = dtype.type(time_step)
  params = (mean_reversion, theta, volvol, rho)
  all_times, keep_mask = prepare_grid(times, time_step, dtype, params, times_grid)
  k = times.shape[0]
  dt = all_times[1:] - all_times[:-1]
  tp = all_times + dt.min() / 2
  kap, th, vv, rh = (_params_at(p, tp, dtype) for p in params)
  steps = dt.shape[0]
  if normal_draws is None:
    normal_draws = draws_lib.generate_mc_normal_draws(
        2, steps, num_samples,
        draws_lib.RandomType.PSEUDO if random_type is None else random_type,
        dtype=dtype, seed=seed, skip=skip, path_range=path_range)
    num_samples = normal_draws.shape[1]
  else:
    normal_draws = np.transpose(np.asarray(normal_draws, dtype), [1, 0, 2])
    num_samples = normal_draws.shape[1]
  x = x0[..., 0] + np.zeros(num_samples, dtype)
  v = x0[..., 1] + np.zeros(num_samples, dtype)
  record = k != 1
  xs, vs = [None] * k, [None] * k
  if record:
    xs[0], vs[0] = x, v
  written = int(keep_mask[0])
  i = 0
  xmax, xmin = np.array(x), np.array(x)      # oracle extension (return_extrema)
  while i < steps and written < k:
    z = normal_draws[i]
    if dt[i] > tolerance:
      v_next = _update_variance(kap[i], th[i], vv[i], v, dt[i], z[..., 0]).astype(dtype)
      x = _update_log_spot(kap[i], th[i], vv[i], rh[i], v, v_next, x, dt[i],
                           z[..., 1]).astype(dtype)
      v = v_next
    xmax, xmin = np.maximum(xmax, x), np.minimum(xmin, x)
    if record:
      xs[written], vs[written] = x, v
    written += int(keep_mask[i + 1])
    i += 1
  if not record:
    out = np.stack([x[:, None], v[:, None]], -1)
    return (out, xmax, xmin) if return_extrema else out
  zeros = np.zeros(num_samples, dtype)
  xs = [zeros if a is None else a for a in xs]
  vs = [zeros if a is None else a for a in vs]
  out = np.stack([np.stack(xs, 0).T, np.stack(vs, 0).T], -1)
  return (out, xmax, xmin) if return_extrema else out
