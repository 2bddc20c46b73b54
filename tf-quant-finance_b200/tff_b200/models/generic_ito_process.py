"""`GenericItoProcess` (`tf_quant_finance/models/generic_ito_process.py`).

`sample_paths` forwards to `euler_sampling.sample` exactly like the reference
(lines 297-317; note that it does not forward `validate_args` there either).
"""
import numpy as np

from tff_b200 import _tensor
from tff_b200.models import euler_sampling
from tff_b200.models import ito_process


class GenericItoProcess(ito_process.ItoProcess):
  """An Ito process defined by (dim, drift_fn, volatility_fn)."""

  def __init__(self, dim, drift_fn, volatility_fn, dtype=None, name=None):
    if dim < 1:
      raise ValueError('Dimension must be 1 or greater.')
    if drift_fn is None or volatility_fn is None:
      raise ValueError('Both drift and volatility functions must be supplied.')
    self._dim = dim
    self._drift_fn = drift_fn
    self._volatility_fn = volatility_fn
    self._dtype = None if dtype is None else _tensor.np_dtype(dtype)
    self._name = name or 'generic_ito_process'

  def dim(self):
    return self._dim

  def dtype(self):
    return self._dtype

  def name(self):
    return self._name

  def drift_fn(self):
    return self._drift_fn

  def volatility_fn(self):
    return self._volatility_fn

  def _sampling_args(self, times, num_samples, initial_state, random_type,
                     seed, time_step, num_time_steps, skip, times_grid,
                     normal_draws, watch_params):
    return dict(
        dim=self._dim, drift_fn=self._drift_fn,
        volatility_fn=self._volatility_fn, times=times,
        num_samples=num_samples, initial_state=initial_state,
        random_type=random_type, time_step=time_step,
        num_time_steps=num_time_steps, seed=seed, skip=skip,
        times_grid=times_grid, normal_draws=normal_draws,
        watch_params=watch_params, dtype=self._dtype)

  def sample_paths(self,
                   times,
                   num_samples=1,
                   initial_state=None,
                   random_type=None,
                   seed=None,
                   swap_memory=True,
                   name=None,
                   time_step=None,
                   num_time_steps=None,
                   skip=0,
                   precompute_normal_draws=True,
                   times_grid=None,
                   normal_draws=None,
                   watch_params=None,
                   validate_args=False):
    """Euler paths `[num_samples, k, dim]` (`generic_ito_process.py:187-317`)."""
    del validate_args  # the reference does not forward it either
    return euler_sampling.sample(
        swap_memory=swap_memory,
        precompute_normal_draws=precompute_normal_draws,
        name=name or (self._name + '_sample_path'),
        **self._sampling_args(times, num_samples, initial_state, random_type,
                              seed, time_step, num_time_steps, skip, times_grid,
                              normal_draws, watch_params))

  def price(self, times, payoffs, num_samples=1, initial_state=None,
            random_type=None, seed=None, time_step=None, num_time_steps=None,
            skip=0, times_grid=None, normal_draws=None, return_stats=False):
    """Fused Euler simulation + payoff reduction (engine extension)."""
    args = self._sampling_args(times, num_samples, initial_state, random_type,
                               seed, time_step, num_time_steps, skip,
                               times_grid, normal_draws, None)
    args.pop('watch_params')
    return euler_sampling.price(payoffs=payoffs, return_stats=return_stats,
                                **args)
