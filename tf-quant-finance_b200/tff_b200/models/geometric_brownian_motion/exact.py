"""Exact log-normal GBM samplers on the device (what the reference's
`GeometricBrownianMotion.sample_paths` /
`MultivariateGeometricBrownianMotion.sample_paths` execute:
`univariate_geometric_brownian_motion.py:155-317`,
`multivariate_geometric_brownian_motion.py:153-282`).

The grid is exactly the requested times (draws `[k, N, dim]`); the state is
the cumulative log-increment, exponentiated when it is stored.
"""
import numpy as np
import torch

from tff_b200 import _tensor
from tff_b200 import engine
from tff_b200.math import piecewise


def _integrate(p, t0, t1, dtype, square=False):
  if callable(p):
    q = p
    if square:
      q = piecewise.PiecewiseConstantFunc(p.jump_locations(), p.values()**2, dtype=dtype)
    return np.asarray(q.integrate(t0, t1), dtype=dtype).reshape(t0.shape)
  v = np.asarray(p, dtype=dtype).reshape(())
  return ((v * v if square else v) * (t1 - t0)).astype(dtype)


def _finish(plan, k, x0, dtype):
  rec = np.arange(-1, k, dtype=np.int32)           # entry 0 (initial state) not recorded
  positive = bool(np.all(x0 > 0))
  try:
    out = plan.paths(rec, k, exp_transform=True)
  finally:
    plan.close()
  if not positive:
    out = out * torch.as_tensor(x0, device=out.device, dtype=out.dtype)
  return out


def _integrate_batched(p, t0, t1, dtype, square=False):
  """`_integrate_parameter` (`univariate_...py:239-259`) with the reference's broadcasting:
  `batch_shape + [k]` for parameters of shape `batch_shape + [1]` or batched piecewise functions."""
  if callable(p):
    q = p
    if square:
      q = piecewise.PiecewiseConstantFunc(p.jump_locations(), p.values()**2, dtype=dtype)
    return np.asarray(q.integrate(t0, t1), dtype=dtype)
  v = np.asarray(p, dtype=dtype)
  return ((v * v if square else v) * (t1 - t0)).astype(dtype)


def _sample_paths_univariate_batched(model, times, x0, num_samples, random_type, seed, skip,
                                     normal_draws):
  """A batch of GBMs (`univariate_...py:261-317`): parameters `batch_shape + [1]`, times `[k]` or
  `batch_shape + [k]`, initial state broadcastable to `batch_shape + [1]`.  The reference draws ONE
  `[k, N]` set of normals without a batch shape (`:277-282`) and broadcasts it: every element of the
  batch runs on the same draws, so each is one launch of the same plan shape with its own per-step
  constants.  Returns `batch_shape + [N, k, 1]`."""
  dt_ = model.dtype()
  k = times.shape[-1]
  all_times = np.concatenate([np.zeros(times.shape[:-1] + (1,), dt_), times], -1)
  mean_int = _integrate_batched(model._mean, all_times[..., :-1], all_times[..., 1:], dt_)
  vol2_int = _integrate_batched(model._volatility, all_times[..., :-1], all_times[..., 1:], dt_, square=True)
  x0 = x0.reshape(1) if x0.ndim == 0 else x0
  if x0.shape[-1] != 1:
    raise ValueError('`initial_state` must be broadcastable to `batch_shape + [1]`')
  batch_shape = np.broadcast_shapes(mean_int.shape[:-1], vol2_int.shape[:-1], x0.shape[:-1],
                                    all_times.shape[:-1])
  mean_int = np.broadcast_to(mean_int, batch_shape + (k,))
  vol2_int = np.broadcast_to(vol2_int, batch_shape + (k,))
  x0 = np.broadcast_to(x0, batch_shape + (1,))
  all_times = np.broadcast_to(all_times, batch_shape + (k + 1,))
  if normal_draws is not None:
    normal_draws = _tensor.from_dlpack(normal_draws)
    if int(normal_draws.shape[2]) != 1:
      raise ValueError('`dim` should be equal to `1` but is {0}'.format(int(normal_draws.shape[2])))
    num_samples = int(normal_draws.shape[0])
  outs = []
  for index in np.ndindex(*batch_shape):
    drift = (mean_int[index] - vol2_int[index] / 2).astype(dt_)
    with np.errstate(invalid='ignore'):
      vol = np.where(vol2_int[index] > 0, np.sqrt(np.maximum(vol2_int[index], 0)), 0).astype(dt_)
    spec = engine.LinearSpec1F(lambda t, d, drift=drift, vol=vol: (np.ones(k, d), drift, vol))
    rng = engine.RngSpec(random_type, seed, skip, normal_draws)
    x0_b = np.asarray(x0[index], dtype=dt_)
    start = np.log(x0_b) if bool(np.all(x0_b > 0)) else np.zeros(1, dt_)
    plan = engine.Plan(spec, np.ascontiguousarray(all_times[index]), k, start.astype(dt_), rng,
                       int(num_samples), dt_)
    outs.append(_finish(plan, k, x0_b, dt_).contiguous())
  out = torch.stack(outs, dim=0)
  return out.reshape(tuple(batch_shape) + tuple(out.shape[1:]))


def sample_paths_univariate(model, times, initial_state=None, num_samples=1,
                            random_type=None, seed=None, skip=0, normal_draws=None):
  dt_ = model.dtype()
  times = _tensor.to_numpy(times, dt_)
  x0_full = np.ones(1, dt_) if initial_state is None else _tensor.to_numpy(initial_state, dt_)

  def batched(p):
    if callable(p):
      return np.ndim(p.jump_locations()) > 1
    return np.ndim(p) > 1 or (np.ndim(p) == 1 and np.shape(p)[0] != 1)
  if times.ndim > 1 or x0_full.ndim > 1 or batched(model._mean) or batched(model._volatility):
    return _sample_paths_univariate_batched(model, times, x0_full, num_samples, random_type, seed, skip,
                                            normal_draws)
  times = times.reshape(-1)
  k = times.shape[0]
  x0 = np.ones(1, dt_) if initial_state is None else _tensor.to_numpy(initial_state, dt_).reshape(-1)
  if x0.shape[0] != 1:
    raise NotImplementedError('batched initial states are not implemented by the B200 engine yet')
  all_times = np.concatenate([np.zeros(1, dt_), times])
  mean_int = _integrate(model._mean, all_times[:-1], all_times[1:], dt_)
  vol2_int = _integrate(model._volatility, all_times[:-1], all_times[1:], dt_, square=True)
  drift = (mean_int - vol2_int / 2).astype(dt_)
  with np.errstate(invalid='ignore'):
    vol = np.where(vol2_int > 0, np.sqrt(np.maximum(vol2_int, 0)), 0).astype(dt_)   # _sqrt_no_nan

  spec = engine.LinearSpec1F(lambda t, d: (np.ones(k, d), drift, vol))
  if normal_draws is not None:
    normal_draws = _tensor.from_dlpack(normal_draws)
    if int(normal_draws.shape[2]) != 1:
      raise ValueError('`dim` should be equal to `1` but is {0}'.format(int(normal_draws.shape[2])))
    num_samples = int(normal_draws.shape[0])
  rng = engine.RngSpec(random_type, seed, skip, normal_draws)
  positive = bool(np.all(x0 > 0))
  start = np.log(x0) if positive else np.zeros(1, dt_)
  plan = engine.Plan(spec, all_times, k, start.astype(dt_), rng, int(num_samples), dt_)
  return _finish(plan, k, x0, dt_)


def sample_paths_multivariate(model, times, initial_state=None, num_samples=1,
                              random_type=None, seed=None, skip=0, normal_draws=None):
  if normal_draws is not None:
    raise NotImplementedError('normal_draws= is not implemented for the multivariate sampler yet')
  dt_ = model.dtype()
  dim = model.dim()
  times = _tensor.to_numpy(times, dt_).reshape(-1)
  k = times.shape[0]
  x0 = (np.ones(dim, dt_) if initial_state is None
        else np.broadcast_to(_tensor.to_numpy(initial_state, dt_).reshape(-1), (dim,)).copy())
  all_times = np.concatenate([np.zeros(1, dt_), times])
  spec = engine.MvGbmSpec(model._means, model._vols, model._corr_matrix, dim, exact_log=True)
  rng = engine.RngSpec(random_type, seed, skip, None)
  positive = bool(np.all(x0 > 0))
  start = np.log(x0) if positive else np.zeros(dim, dt_)
  plan = engine.Plan(spec, all_times, k, start.astype(dt_), rng, int(num_samples), dt_)
  return _finish(plan, k, x0, dt_)
