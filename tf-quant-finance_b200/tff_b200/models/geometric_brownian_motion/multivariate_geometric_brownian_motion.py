"""Multivariate geometric Brownian motion
(`models/geometric_brownian_motion/multivariate_geometric_brownian_motion.py`).

  dX_i = means_i X_i dt + volatilities_i X_i dW_i,   corr(dW_i, dW_j) = corr_matrix_ij

`drift_fn()` / `volatility_fn()` (reference lines 130-151) feed the Euler
engine (`sample_paths_euler`, `price_euler`): the Cholesky factor is computed
once on the host instead of once per step per path.
"""
import numpy as np

from tff_b200 import _tensor
from tff_b200.models import closures
from tff_b200.models import euler_sampling
from tff_b200.models import ito_process


class MultivariateGeometricBrownianMotion(ito_process.ItoProcess):
  """Multivariate Geometric Brownian Motion."""

  def __init__(self, dim, means=0.0, volatilities=1.0, corr_matrix=None,
               dtype=None, name=None):
    self._name = name or 'multivariate_geometric_brownian_motion'
    self._dtype = _tensor.infer_dtype(means, dtype, default=np.float32)
    self._dim = int(dim)
    self._means = np.broadcast_to(_tensor.to_numpy(means, self._dtype), (self._dim,)).copy()
    self._vols = np.broadcast_to(_tensor.to_numpy(volatilities, self._dtype), (self._dim,)).copy()
    if corr_matrix is None:
      self._corr_matrix = None
    else:
      self._corr_matrix = _tensor.to_numpy(corr_matrix, self._dtype)
      if list(self._corr_matrix.shape) != [self._dim, self._dim]:
        raise ValueError('`corr_matrix` must be of shape [{0}, {0}] but is '
                         'of shape {1}'.format(self._dim, list(self._corr_matrix.shape)))
    self._drift_fn, self._vol_fn = closures.mvgbm_closures(
        self._means, self._vols, self._corr_matrix, self._dim)

  def dim(self):
    return self._dim

  def dtype(self):
    return self._dtype

  def name(self):
    return self._name

  def drift_fn(self):
    return self._drift_fn

  def volatility_fn(self):
    return self._vol_fn

  def _euler_args(self, times, initial_state, num_samples, random_type, seed, skip,
                  time_step, num_time_steps, times_grid):
    if initial_state is None:
      initial_state = np.ones(self._dim, dtype=self._dtype)
    return dict(dim=self._dim, drift_fn=self._drift_fn, volatility_fn=self._vol_fn,
                times=times, time_step=time_step, num_time_steps=num_time_steps,
                num_samples=num_samples, initial_state=initial_state,
                random_type=random_type, seed=seed, skip=skip, times_grid=times_grid,
                dtype=self._dtype)

  def sample_paths_euler(self, times, initial_state=None, num_samples=1,
                         random_type=None, seed=None, skip=0, time_step=None,
                         num_time_steps=None, times_grid=None):
    """Euler-Maruyama paths `[num_samples, k, dim]` through the closures."""
    return euler_sampling.sample(**self._euler_args(
        times, initial_state, num_samples, random_type, seed, skip, time_step,
        num_time_steps, times_grid))

  def price_euler(self, times, payoffs, initial_state=None, num_samples=1,
                  random_type=None, seed=None, skip=0, time_step=None,
                  num_time_steps=None, times_grid=None, return_stats=False):
    """Fused Euler simulation + payoff reduction (component -1 = basket mean)."""
    return euler_sampling.price(payoffs=payoffs, return_stats=return_stats,
                                **self._euler_args(times, initial_state, num_samples,
                                                   random_type, seed, skip, time_step,
                                                   num_time_steps, times_grid))

  def sample_paths(self, times, initial_state=None, num_samples=1,
                   random_type=None, seed=None, skip=0, normal_draws=None,
                   name=None):
    """Exact log-normal sampler (`multivariate_...py:153-282`)."""
    from tff_b200.models.geometric_brownian_motion import exact  # pylint: disable=g-import-not-at-top
    del name
    return exact.sample_paths_multivariate(
        self, times, initial_state=initial_state, num_samples=num_samples,
        random_type=random_type, seed=seed, skip=skip, normal_draws=normal_draws)
