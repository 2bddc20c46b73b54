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


def sample_paths_univariate(model, times, initial_state=None, num_samples=1,
                            random_type=None, seed=None, skip=0, normal_draws=None):
  dt_ = model.dtype()
  times = _tensor.to_numpy(times, dt_).reshape(-1)
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
