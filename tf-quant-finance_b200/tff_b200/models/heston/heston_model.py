"""Heston model (`tf_quant_finance/models/heston/heston_model.py`).

  dX = -V/2 dt + sqrt(V) dW_X,   dV = kappa (theta - V) dt + volvol sqrt(V) dW_V

* `drift_fn()` / `volatility_fn()` are the Euler closures of the reference
  (lines 143-173); pushed through `GenericItoProcess.sample_paths` /
  `euler_sampling.sample` they give the Euler-Maruyama scheme
  (`sample_paths_euler`, `price_euler`).
* `sample_paths` is Andersen's Quadratic-Exponential scheme exactly as the
  reference's `HestonModel.sample_paths` (lines 177-460, 522-639).
"""
import numpy as np

from tff_b200 import _tensor
from tff_b200.math import piecewise
from tff_b200.models import closures
from tff_b200.models import generic_ito_process


def _convert(param, dtype):
  if isinstance(param, piecewise.PiecewiseConstantFunc):
    return param
  if getattr(param, 'is_piecewise_constant', False):
    return param
  return _tensor.to_numpy(param, dtype)


class HestonModel(generic_ito_process.GenericItoProcess):
  """Heston Model with piecewise constant parameters."""

  def __init__(self, mean_reversion, theta, volvol, rho, dtype=None, name=None):
    self._name = name or 'heston_model'
    dt = _tensor.np_dtype(dtype, np.float32)
    self._mean_reversion = _convert(mean_reversion, dt)
    self._theta = _convert(theta, dt)
    self._volvol = _convert(volvol, dt)
    self._rho = _convert(rho, dt)
    drift_fn, vol_fn = closures.heston_closures(
        self._mean_reversion, self._theta, self._volvol, self._rho)
    super().__init__(2, drift_fn, vol_fn, dt, self._name)

  def sample_paths_euler(self, times, initial_state, num_samples=1,
                         random_type=None, seed=None, time_step=None, skip=0,
                         num_time_steps=None, times_grid=None,
                         normal_draws=None):
    """Euler-Maruyama paths `[num_samples, k, 2]` = (log-spot, variance)."""
    return generic_ito_process.GenericItoProcess.sample_paths(
        self, times, num_samples=num_samples, initial_state=initial_state,
        random_type=random_type, seed=seed, time_step=time_step,
        num_time_steps=num_time_steps, skip=skip, times_grid=times_grid,
        normal_draws=normal_draws)

  def sample_paths(self, times, initial_state, num_samples=1, random_type=None,
                   seed=None, time_step=None, skip=0, tolerance=1e-6,
                   num_time_steps=None, precompute_normal_draws=True,
                   times_grid=None, normal_draws=None, name=None):
    """Andersen QE paths (`heston_model.py:177-460`)."""
    from tff_b200.models.heston import qe  # pylint: disable=g-import-not-at-top
    del precompute_normal_draws, name
    return qe.sample_paths(
        self, times, initial_state, num_samples=num_samples,
        random_type=random_type, seed=seed, time_step=time_step, skip=skip,
        tolerance=tolerance, num_time_steps=num_time_steps,
        times_grid=times_grid, normal_draws=normal_draws)

  def price(self, times, payoffs, num_samples=1, initial_state=None,
            random_type=None, seed=None, time_step=None, num_time_steps=None,
            skip=0, times_grid=None, normal_draws=None, return_stats=False,
            scheme='euler', tolerance=1e-6):
    """Fused simulation + payoff reduction (engine extension; nothing stored).
    `scheme='euler'`: the Euler closures, as `GenericItoProcess.price`;
    `scheme='qe'`: the QE scheme of `sample_paths` (what the reference's
    `HestonModel.sample_paths` runs)."""
    if scheme == 'euler':
      return generic_ito_process.GenericItoProcess.price(
          self, times, payoffs, num_samples=num_samples, initial_state=initial_state,
          random_type=random_type, seed=seed, time_step=time_step,
          num_time_steps=num_time_steps, skip=skip, times_grid=times_grid,
          normal_draws=normal_draws, return_stats=return_stats)
    if scheme != 'qe':
      raise ValueError("scheme must be 'euler' or 'qe'")
    from tff_b200.models.heston import qe  # pylint: disable=g-import-not-at-top
    return qe.price(self, times, payoffs, initial_state, num_samples=num_samples,
                    random_type=random_type, seed=seed, time_step=time_step, skip=skip,
                    tolerance=tolerance, num_time_steps=num_time_steps,
                    times_grid=times_grid, normal_draws=normal_draws,
                    return_stats=return_stats)

  def expected_total_variance(self, future_times, initial_var, name=None):
    """`heston_model.py:462-509` (host, numpy)."""
    del name
    for pname in ('_mean_reversion', '_theta'):
      if callable(getattr(self, pname)):
        raise ValueError(f'Only constant values supported for {pname}')
    t = _tensor.to_numpy(future_times, self._dtype)
    v0 = _tensor.to_numpy(initial_var, self._dtype)
    k, th = self._mean_reversion, self._theta
    return (v0 - th) * (1 - np.exp(-k * t)) / k + th * t
