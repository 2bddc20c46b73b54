"""The Euler sampling method for Ito processes -- B200 engine.

Drop-in for `tf_quant_finance.models.euler_sampling.sample`
(`models/euler_sampling.py:27-332`): same arguments, same output shape
`batch_shape + [num_samples, k, dim]`, same draw layout for every
`random_type`, so results equal the reference's on identical seeds.  The
draws tensor of the reference is never built: normals are generated in the
kernel that steps the paths.  `price` is the fused extension that also
reduces payoffs in-kernel (no path leaves the registers).
"""
import collections

import numpy as np
import torch

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import distributed
from tff_b200 import engine
from tff_b200.math import random
from tff_b200.models import closures
from tff_b200.models import utils


class InvalidArgumentError(ValueError):
  """Stands in for `tf.errors.InvalidArgumentError` (TensorFlow is optional)."""


def _prepare(dim, drift_fn, volatility_fn, times, time_step, num_time_steps,
             num_samples, initial_state, random_type, seed, skip, times_grid,
             normal_draws, watch_params, validate_args, tolerance, dtype, use_cache=False):
  """Argument normalisation of `sample` (`euler_sampling.py:232-310`)."""
  # `watch_params` (euler_sampling.py:393-402, 467-510) only chooses how TensorFlow
  # DIFFERENTIATES the loop (custom_loops.for_loop instead of tf.while_loop); the sampled
  # paths do not depend on it, and they are what this engine returns.  Sensitivities are
  # carried in-kernel by the tangent closures instead (closures.affine_tangent_closures,
  # closures.heston_tangent_closures: dX/dX0, dX/dtheta as extra state components).
  del watch_params
  dtype = _tensor.infer_dtype(times, dtype)
  times = _tensor.to_numpy(times, dtype).reshape(-1)
  if tolerance is None:
    tolerance = 1e-10 if dtype == np.float64 else 1e-6
  tolerance = dtype.type(tolerance)
  if validate_args and not np.all(times[1:] > times[:-1] + tolerance):
    raise InvalidArgumentError('`times` increments should be greater '
                               'than tolerance {0}'.format(tolerance))
  if initial_state is None:
    initial_state = np.zeros(dim, dtype=dtype)
  initial_state = _tensor.to_numpy(initial_state, dtype)
  batch_shape = initial_state.shape[:-2]
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
      num_time_steps = int(num_time_steps)
      time_step = dtype.type(times[-1] / dtype.type(num_time_steps))
    else:
      time_step = dtype.type(_tensor.to_numpy(time_step))
  else:
    times_grid = _tensor.to_numpy(times_grid, dtype)
    if validate_args and not np.all(
        times_grid[1:] > times_grid[:-1] + tolerance):
      raise InvalidArgumentError('`times_grid` increments should be greater '
                                 'than tolerance {0}'.format(tolerance))
  all_times, keep_mask, _ = utils.prepare_grid(
      times=times, time_step=time_step, num_time_steps=num_time_steps,
      times_grid=times_grid, tolerance=tolerance, dtype=dtype)

  if (normal_draws is None and random_type is not None
      and random_type.value in (random.RandomType.HALTON.value,
                                random.RandomType.HALTON_RANDOMIZED.value)):
    # The Halton sequences have no in-kernel generator: their normals
    # are materialised like the reference does for every random type
    # (euler_sampling.py:365-374, utils.py:20-128) and fed as `normal_draws`.
    if batch_shape:
      raise NotImplementedError('batched processes with HALTON draws are not implemented yet')
    draws = utils.generate_mc_normal_draws(
        num_normal_draws=dim, num_time_steps=all_times.shape[0] - 1,
        num_sample_paths=int(num_samples), random_type=random_type, skip=skip, seed=seed,
        dtype=dtype)
    normal_draws = draws.permute(1, 0, 2).contiguous()           # [N, steps, dim]
  if normal_draws is not None:
    normal_draws = _tensor.from_dlpack(normal_draws)
    # batch_shape + [num_samples, num_time_points, dim]
    num_samples = int(normal_draws.shape[-3])
    draws_dim = int(normal_draws.shape[-1])
    if dim != draws_dim:
      raise ValueError(
          '`dim` should be equal to `normal_draws.shape[2]` but are '
          '{0} and {1} respectively'.format(dim, draws_dim))
    if validate_args and int(normal_draws.shape[-2]) != keep_mask.shape[0] - 1:
      raise InvalidArgumentError('`num_time_steps` should be equal to '
                                 '`tf.shape(normal_draws)[1]`')
  if normal_draws is not None and (batch_shape or normal_draws.dim() > 3):
    raise NotImplementedError(
        'batched `normal_draws` are not implemented by the B200 engine yet.')
  spec = closures.resolve_spec(drift_fn, volatility_fn, dim)
  if isinstance(spec, engine.ProbedAffineSpec):
    spec.initial_state_hint = np.asarray(initial_state, dtype=np.float64).reshape(-1, dim)[0]
  if getattr(spec, 'user_dim', spec.dim) != dim:
    raise ValueError('`dim` is {} but the model has dimension {}'.format(
        dim, getattr(spec, 'user_dim', spec.dim)))
  if hasattr(spec, 'extend_initial_state') and (batch_shape or normal_draws is not None):
    raise NotImplementedError(
        'tangent-carrying closures support neither batched processes nor normal_draws yet')
  num_steps, record_slot = engine.record_plan(keep_mask, times.shape[0])
  num_samples = int(num_samples)
  # one plan per element of the batch of processes (a single one without batch)
  x0_full = np.broadcast_to(initial_state, batch_shape + initial_state.shape[len(batch_shape):])
  batch_size = int(np.prod(batch_shape)) if batch_shape else 1
  antithetic = normal_draws is None and random_type is not None and random_type.value in (
      random.RandomType.PSEUDO_ANTITHETIC.value, random.RandomType.STATELESS_ANTITHETIC.value)
  plans = []
  for bi, index in enumerate(np.ndindex(*batch_shape)):
    x0 = np.asarray(x0_full[index]).reshape(-1, dim)
    x0_paths = None
    if x0.shape[0] != 1 and not np.all(x0 == x0[0]):
      # one initial state per path (`initial_state` of shape [num_samples, dim])
      if x0.shape[0] != int(num_samples):
        raise ValueError('`initial_state` has {} rows but `num_samples` is {}'.format(
            x0.shape[0], int(num_samples)))
      if hasattr(spec, 'extend_initial_state') or spec.kind == _lib.MODEL_MVGBM:
        raise NotImplementedError(
            'per-path initial states are not implemented for this model yet.')
      x0_paths = x0
    spec_b = spec.for_batch(index, batch_shape) if (
        batch_shape and hasattr(spec, 'for_batch')) else spec
    # draw units of a batch (models/utils.py:98-107): [batch, N] row-major, or
    # [N/2, batch] for the antithetic types (sample_shape = [N] + batch_shape)
    stride, offset = (batch_size, bi) if antithetic else (1, bi * num_samples)
    rng = engine.RngSpec(random_type, seed, skip, normal_draws, unit_stride=stride,
                         unit_offset=offset)
    x0_plan = spec_b.extend_initial_state(x0[0]) if hasattr(spec_b, 'extend_initial_state') else x0[0]
    if use_cache and x0_paths is None:
      plans.append(engine.cached_plan(spec_b, all_times, num_steps, x0_plan, rng, num_samples, dtype))
    else:
      plans.append(engine.Plan(spec_b, all_times, num_steps, x0_plan, rng, num_samples, dtype,
                               x0_paths=x0_paths))
  return plans, record_slot, times.shape[0], batch_shape


def sample(dim,
           drift_fn,
           volatility_fn,
           times,
           time_step=None,
           num_time_steps=None,
           num_samples=1,
           initial_state=None,
           random_type=None,
           seed=None,
           swap_memory=True,
           skip=0,
           precompute_normal_draws=True,
           times_grid=None,
           normal_draws=None,
           watch_params=None,
           validate_args=False,
           tolerance=None,
           dtype=None,
           name=None):
  """Returns a sample of paths from the process using the Euler method.

  Same contract as the reference's `sample`.  `swap_memory`,
  `precompute_normal_draws` and `name` only steer TensorFlow's execution and
  are ignored: draws are never precomputed, and the result is defined to equal
  the reference's precomputed-draws path.

  Returns:
    CUDA tensor of shape `batch_shape + [num_samples, k, dim]` (a zero-copy view
    of a time-major buffer; call `.contiguous()` for sample-major memory).
  """
  del swap_memory, precompute_normal_draws, name
  plans, record_slot, k, batch_shape = _prepare(
      dim, drift_fn, volatility_fn, times, time_step, num_time_steps,
      num_samples, initial_state, random_type, seed, skip, times_grid,
      normal_draws, watch_params, validate_args, tolerance, dtype)
  try:
    if not batch_shape:
      return plans[0].paths(record_slot, k)
    n = plans[0].num_samples
    buf = _tensor.empty((len(plans), k, dim, n), plans[0].dtype)
    for bi, plan in enumerate(plans):
      plan.paths(record_slot, k, out=buf[bi])
    nb = len(batch_shape)
    out = buf.reshape(tuple(batch_shape) + (k, dim, n))
    return out.permute(*range(nb), nb + 2, nb, nb + 1)
  finally:
    for plan in plans:
      plan.close()


# A pricing call whose arguments are all plain numbers / small arrays is recognised by
# their CONTENT the second time it is made (a calibration loop, a risk run re-pricing the
# same book): the plan, the payoff descriptors and the result buffers stay bound and the
# call is one FFI call.  Models with callable parameters, `normal_draws=`, sharded runs
# and anything large are never bound (`_call_key` returns None).
_CALLS = collections.OrderedDict()
_CALLS_SIZE = 8


def _small(a):
  if a is None:
    return None
  a = np.asarray(_tensor.to_numpy(a))
  if a.size > 64 or a.dtype == object:
    raise OverflowError
  return (a.dtype.str, a.shape, a.tobytes())


def _call_key(dim, drift_fn, volatility_fn, times, payoffs, time_step, num_time_steps,
              num_samples, initial_state, random_type, seed, skip, times_grid, normal_draws,
              tolerance, dtype):
  if normal_draws is not None or distributed._SHARDED is not None:   # pylint: disable=protected-access
    return None
  spec = getattr(drift_fn, 'tqf_spec', None)
  if spec is None or spec is not getattr(volatility_fn, 'tqf_spec', None):
    return None
  model = spec.constant_key()
  if model is None or not all(isinstance(p, engine.Payoff) for p in payoffs):
    return None
  try:
    return (torch.cuda.current_device(), int(dim), model, _small(times), tuple(p.key() for p in payoffs),
            _small(time_step), None if num_time_steps is None else int(num_time_steps),
            int(num_samples), _small(initial_state),
            None if random_type is None else random_type.value, _small(seed), int(skip),
            _small(times_grid), tolerance, None if dtype is None else np.dtype(_tensor.np_dtype(dtype)).str)
  except (OverflowError, TypeError, ValueError):
    return None


def price(dim, drift_fn, volatility_fn, times, payoffs, time_step=None,
          num_time_steps=None, num_samples=1, initial_state=None,
          random_type=None, seed=None, skip=0, times_grid=None,
          normal_draws=None, validate_args=False, tolerance=None, dtype=None,
          return_stats=False):
  """Fused mode: Monte-Carlo means of `payoffs` on the state at `times[-1]`.

  Equals `mean(payoff(sample(...)[:, -1, :]))` of the materialising path (and
  barrier payoffs monitored on every grid point) without storing any path.
  Returns a float64 numpy array `[len(payoffs)]`; with `return_stats` also the
  standard errors and the number of non-finite payoffs.
  """
  key = _call_key(dim, drift_fn, volatility_fn, times, payoffs, time_step, num_time_steps,
                  num_samples, initial_state, random_type, seed, skip, times_grid, normal_draws,
                  tolerance, dtype)
  bound = _CALLS.get(key) if key is not None else None
  if bound is not None and bound.alive():
    _CALLS.move_to_end(key)
    plan, sums = bound.plan, bound.sums()
  else:
    plans, _, _, batch_shape = _prepare(
        dim, drift_fn, volatility_fn, times, time_step, num_time_steps,
        num_samples, initial_state, random_type, seed, skip, times_grid,
        normal_draws, None, validate_args, tolerance, dtype, use_cache=True)
    if batch_shape:
      for plan in plans:
        plan.release()
      raise NotImplementedError('batched processes are not supported by `price` yet')
    plan = plans[0]
    try:
      sums = distributed.price_sums_host(plan, payoffs)
      if key is not None and plan.cached:
        _CALLS[key] = engine.HostPricing(plan, payoffs)
        while len(_CALLS) > _CALLS_SIZE:
          _CALLS.popitem(last=False)
    finally:
      plan.release()
  n = float(plan.num_samples)
  mean = sums[:, 0] / n
  if not return_stats:
    return mean
  var = np.maximum(sums[:, 1] / n - mean**2, 0.0)
  return mean, np.sqrt(var / n), sums[:, 2].copy()


__all__ = ['sample', 'price']
