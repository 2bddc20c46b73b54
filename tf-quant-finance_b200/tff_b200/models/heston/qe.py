"""Andersen's Quadratic-Exponential scheme for the Heston model on the device:
what the reference's `HestonModel.sample_paths` executes
(`models/heston/heston_model.py:177-460, 522-639`).

Host side: the Heston-specific time grid (`_prepare_grid` 575-639 -- uniform
grid, requested times and parameter jumps merged by a STABLE argsort with
duplicates kept; zero-length steps are no-ops that still consume a row of
draws) and the per-step constants; device side: `HestonQeModel` in the fused
path kernel.
"""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import distributed
from tff_b200 import engine
from tff_b200.math import piecewise
from tff_b200.models import utils


def _params_at(p, t, dtype):
  if callable(p):
    return np.asarray(p(t), dtype=dtype).reshape(t.shape)
  return np.asarray(p, dtype=dtype) * np.ones_like(t)


def prepare_grid(times, time_step, dtype, params, times_grid=None):
  """`_prepare_grid` (`heston_model.py:575-639`) -> (all_times, mask)."""
  jumps = [np.asarray(p.jump_locations(), dtype=dtype).reshape(-1) for p in params
           if isinstance(p, piecewise.PiecewiseConstantFunc)]
  if times_grid is None:
    grid = utils._tf_range(0.0, times[-1], time_step, dtype)
    all_times = np.concatenate([grid, times] + jumps).astype(dtype)
    mask = np.concatenate([np.zeros(grid.shape, bool), np.ones(times.shape, bool)] +
                          [np.zeros(j.shape, bool) for j in jumps])
    perm = np.argsort(all_times, kind='stable')
    return all_times[perm], mask[perm]
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=time_step,
                                          times_grid=times_grid, dtype=dtype)
  return all_times, mask


class HestonQeSpec(engine.ModelSpec):
  """Per-step constants of the QE step (parameters at `all_times + min(dt)/2`,
  index i for step i, as `_sample_paths` 337-340 does)."""
  kind, dim, num_factors, num_coef = _lib.MODEL_HESTON_QE, 2, 2, 10

  def __init__(self, mean_reversion, theta, volvol, rho, tolerance):
    self.params = (mean_reversion, theta, volvol, rho)
    self.tolerance = tolerance

  def coef_table(self, all_times, dtype):
    t = np.asarray(all_times, dtype=dtype)
    dt = t[1:] - t[:-1]
    tp = t + dt.min() / 2 if dt.shape[0] else t
    kap, th, vv, rho = (_params_at(p, tp, dtype)[:-1] for p in self.params)
    e = np.exp(-kap * dt)
    vv2 = vv**2
    with np.errstate(all='ignore'):
      c_s1 = vv2 * e / kap * (1 - e)
      c_s0 = th * vv2 / 2 / kap * (1 - e)**2
      k0 = -rho * kap * th / vv * dt
      k1 = 0.5 * dt * (kap * rho / vv - 0.5) - rho / vv
      k2 = 0.5 * dt * (kap * rho / vv - 0.5) + rho / vv
      k3 = 0.5 * dt * (1 - rho**2)
      k4 = 0.5 * dt * (1 - rho**2)
    active = (dt > np.dtype(dtype).type(self.tolerance)).astype(dtype)
    cols = [active, e, th, c_s1, c_s0, k0, k1, k2, k3, k4]
    return np.nan_to_num(np.stack(cols, -1).astype(np.float64))


def _plan(model, times, initial_state, num_samples, random_type, seed, time_step, skip,
          tolerance, num_time_steps, times_grid, normal_draws, use_cache=False):
  """(plan, record_slot, k) of one QE sampling call (`heston_model.py:177-340`)."""
  dt_ = model.dtype()
  times = _tensor.to_numpy(times, dt_).reshape(-1)
  x0 = _tensor.to_numpy(initial_state, dt_)
  if x0.reshape(-1).shape[0] != 2:
    raise NotImplementedError('per-path / batched initial states are not '
                              'implemented by the B200 engine yet')
  if times_grid is None:
    if time_step is None:
      if num_time_steps is None:
        raise ValueError(
            'When `times_grid` is not supplied, either `num_time_steps` '
            'or `time_step` should be defined.')
      time_step = dt_.type(times[-1] / dt_.type(int(num_time_steps)))
    else:
      if num_time_steps is not None:
        raise ValueError(
            'Both `time_step` and `num_time_steps` can not be `None` '
            'simultaneously when calling sample_paths of HestonModel.')
      time_step = dt_.type(_tensor.to_numpy(time_step))
  else:
    times_grid = _tensor.to_numpy(times_grid, dt_)
  params = (model._mean_reversion, model._theta, model._volvol, model._rho)
  all_times, mask = prepare_grid(times, time_step, dt_, params, times_grid)
  if normal_draws is not None:
    normal_draws = _tensor.from_dlpack(normal_draws)
    num_samples = int(normal_draws.shape[0])
  num_steps, record_slot = engine.record_plan(mask, times.shape[0])
  spec = HestonQeSpec(*params, tolerance)
  rng = engine.RngSpec(random_type, seed, skip, normal_draws)
  make = engine.cached_plan if use_cache else engine.Plan
  plan = make(spec, all_times, num_steps, x0.reshape(-1), rng, int(num_samples), dt_)
  return plan, record_slot, times.shape[0]


def sample_paths(model, times, initial_state, num_samples=1, random_type=None,
                 seed=None, time_step=None, skip=0, tolerance=1e-6,
                 num_time_steps=None, times_grid=None, normal_draws=None):
  """`[num_samples, k, 2]` = (log-spot, variance) QE paths on the device."""
  plan, record_slot, k = _plan(model, times, initial_state, num_samples, random_type, seed,
                               time_step, skip, tolerance, num_time_steps, times_grid,
                               normal_draws)
  try:
    return plan.paths(record_slot, k)
  finally:
    plan.close()


def price(model, times, payoffs, initial_state, num_samples=1, random_type=None,
          seed=None, time_step=None, skip=0, tolerance=1e-6, num_time_steps=None,
          times_grid=None, normal_draws=None, return_stats=False):
  """Fused mode of the QE scheme: Monte-Carlo means of `payoffs` on the state at
  `times[-1]` of `sample_paths(...)` with the same arguments (barrier payoffs
  monitored on every grid point), no path stored.  Engine extension -- the
  reference would take `mean(payoff(sample_paths(...)))`."""
  plan, _, _ = _plan(model, times, initial_state, num_samples, random_type, seed,
                     time_step, skip, tolerance, num_time_steps, times_grid, normal_draws,
                     use_cache=True)
  try:
    sums = distributed.price_sums_host(plan, payoffs)
  finally:
    plan.release()
  n = float(plan.num_samples)
  mean = sums[:, 0] / n
  if not return_stats:
    return mean
  var = np.maximum(sums[:, 1] / n - mean**2, 0.0)
  return mean, np.sqrt(var / n), sums[:, 2]
