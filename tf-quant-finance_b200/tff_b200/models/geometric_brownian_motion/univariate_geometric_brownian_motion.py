"""Univariate geometric Brownian motion
(`models/geometric_brownian_motion/univariate_geometric_brownian_motion.py`).

  dX = mean(t) X dt + volatility(t) X dW

`drift_fn()` / `volatility_fn()` (reference lines 127-153) feed the Euler
engine (`sample_paths_euler`, `price_euler`).
"""
import numpy as np

from tff_b200 import _tensor
from tff_b200.math import piecewise
from tff_b200.models import closures
from tff_b200.models import euler_sampling
from tff_b200.models import ito_process


class GeometricBrownianMotion(ito_process.ItoProcess):
  """Geometric Brownian Motion with constant or piecewise constant params."""

  def __init__(self, mean, volatility, dtype=None, name=None):
    self._name = name or 'geometric_brownian_motion'
    dt = None if dtype is None else _tensor.np_dtype(dtype)
    self._mean, self._mean_is_constant = piecewise.convert_to_tensor_or_func(
        mean, dtype=dt)
    if dt is None:
      dt = (self._mean.dtype() if callable(self._mean) else
            _tensor.infer_dtype(mean, None))
    self._dtype = np.dtype(dt)
    (self._volatility,
     self._volatility_is_constant) = piecewise.convert_to_tensor_or_func(
         volatility, dtype=self._dtype)
    if self._mean_is_constant:
      self._mean = np.asarray(self._mean, dtype=self._dtype)
    # parameters of shape `batch_shape + [1]` (or batched PiecewiseConstantFuncs) describe a batch of
    # processes (`univariate_...py:66-80`): `sample_paths` loops over the batch, the Euler closures
    # carry the arrays (`GbmSpec1F.for_batch`)
    self._dim = 1
    self._drift_fn, self._vol_fn = closures.gbm_closures(
        self._mean, self._volatility)

  def dim(self):
    return self._dim

  def dtype(self):
    return self._dtype

  def name(self):
    return self._name

  def drift_is_constant(self):
    return self._mean_is_constant

  def volatility_is_constant(self):
    return self._volatility_is_constant

  def drift_fn(self):
    return self._drift_fn

  def volatility_fn(self):
    return self._vol_fn

  def sample_paths_euler(self, times, initial_state=None, num_samples=1,
                         random_type=None, seed=None, skip=0, time_step=None,
                         num_time_steps=None, times_grid=None,
                         normal_draws=None):
    """Euler-Maruyama paths `[num_samples, k, 1]` through the closures."""
    return euler_sampling.sample(
        1, self._drift_fn, self._vol_fn, times, time_step=time_step,
        num_time_steps=num_time_steps, num_samples=num_samples,
        initial_state=initial_state, random_type=random_type, seed=seed,
        skip=skip, times_grid=times_grid, normal_draws=normal_draws,
        dtype=self._dtype)

  def sample_paths(self, times, initial_state=None, num_samples=1,
                   random_type=None, seed=None, skip=0, normal_draws=None,
                   name=None):
    """Exact log-normal sampler (`univariate_...py:155-317`)."""
    from tff_b200.models.geometric_brownian_motion import exact  # pylint: disable=g-import-not-at-top
    del name
    return exact.sample_paths_univariate(
        self, times, initial_state=initial_state, num_samples=num_samples,
        random_type=random_type, seed=seed, skip=skip,
        normal_draws=normal_draws)
