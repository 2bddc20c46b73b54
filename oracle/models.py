"""Oracle (test infrastructure): drift / volatility closures of the models.

Restates, in numpy, the closures the reference hands to the Euler sampler:
  * GBM            `models/geometric_brownian_motion/univariate_geometric_brownian_motion.py:127-153`
  * MV-GBM         `models/geometric_brownian_motion/multivariate_geometric_brownian_motion.py:130-151`
  * Heston         `models/heston/heston_model.py:143-173`
  * Hull-White     `models/hull_white/vector_hull_white.py:275-306`
and `math/piecewise.py:19-208` (`PiecewiseConstantFunc`, batch-free form).
"""
import numpy as np


class PiecewiseConstantFunc:
  """Left-continuous piecewise constant function (`math/piecewise.py:144-176`).

  f(x) = values[i] for jump_locations[i-1] < x <= jump_locations[i].
  `__call__` is batch-free (1-D `jump_locations`, `values` may carry an event shape); `integrate`
  also takes batched functions.
  """
  is_piecewise_constant = True

  def __init__(self, jump_locations, values, dtype=None):
    self._jumps = np.asarray(jump_locations, dtype=dtype)
    self._dtype = self._jumps.dtype
    self._values = np.asarray(values, dtype=self._dtype)

  def jump_locations(self):
    return self._jumps

  def values(self):
    return self._values

  def dtype(self):
    return self._dtype

  def __call__(self, x, left_continuous=True):
    x = np.asarray(x, dtype=self._dtype)
    side = 'left' if left_continuous else 'right'
    idx = np.searchsorted(self._jumps, x, side=side)
    return self._values[idx]

  def integrate(self, x1, x2):
    """`math/piecewise.py:178-208` for x1 <= x2 (scalar-valued).  A batched function
    (`jump_locations` of shape `batch_shape + [n]`) takes `x1`, `x2` broadcastable to
    `batch_shape + [num_points]` and integrates element by element of the batch."""
    x1 = np.asarray(x1, dtype=self._dtype)
    x2 = np.asarray(x2, dtype=self._dtype)
    if self._jumps.ndim > 1:
      batch_shape = self._jumps.shape[:-1]
      x1, x2 = np.broadcast_arrays(x1, x2)
      x1 = np.broadcast_to(x1, batch_shape + x1.shape[-1:])
      x2 = np.broadcast_to(x2, batch_shape + x2.shape[-1:])
      out = np.empty(x1.shape, dtype=self._dtype)
      for b in np.ndindex(*batch_shape):
        out[b] = PiecewiseConstantFunc(self._jumps[b], self._values[b], dtype=self._dtype).integrate(x1[b], x2[b])
      return out
    knots = self._jumps
    out = np.zeros(np.broadcast(x1, x2).shape, dtype=self._dtype)
    lo = np.concatenate([[-np.inf], knots])
    hi = np.concatenate([knots, [np.inf]])
    for i in range(self._values.shape[0]):
      a = np.maximum(x1, lo[i])
      b = np.minimum(x2, hi[i])
      out = out + self._values[i] * np.maximum(b - a, 0)
    return out


def _param(p, t, dtype):
  """A scalar parameter or a PiecewiseConstantFunc evaluated at scalar t."""
  if callable(p):
    return np.asarray(p(np.asarray([t], dtype=dtype)), dtype=dtype)[0]
  return np.asarray(p, dtype=dtype)


def gbm_closures(mean, volatility, dtype):
  """Univariate GBM: a = mean(t) x, S = vol(t) x[..., None]."""
  dtype = np.dtype(dtype)

  def drift_fn(t, x):
    return _param(mean, t, dtype) * x

  def vol_fn(t, x):
    return _param(volatility, t, dtype) * x[..., None]
  return drift_fn, vol_fn


def mvgbm_closures(means, volatilities, corr_matrix, dtype):
  """a_i = mu_i x_i; S_ij = sigma_i x_i L_ij, L = cholesky(corr) each step."""
  dtype = np.dtype(dtype)
  means = np.asarray(means, dtype=dtype)
  vols_p = np.asarray(volatilities, dtype=dtype)
  corr = None if corr_matrix is None else np.asarray(corr_matrix, dtype=dtype)

  def drift_fn(t, x):
    del t
    return means * x

  def vol_fn(t, x):
    del t
    vols = vols_p * x
    if corr is not None:
      chol = np.linalg.cholesky(corr).astype(dtype)
      return vols[..., None] * chol
    return vols[..., None] * np.eye(x.shape[-1], dtype=dtype)
  return drift_fn, vol_fn


def heston_closures(mean_reversion, theta, volvol, rho, dtype):
  """State [log-spot X, variance V] (`heston_model.py:143-173`)."""
  dtype = np.dtype(dtype)

  def vol_fn(t, x):
    vol = np.sqrt(np.abs(x[..., 1]))
    zeros = np.zeros_like(vol)
    r = _param(rho, t, dtype)
    vv = _param(volvol, t, dtype)
    col1 = np.stack([vol, vv * r * vol], -1)
    col2 = np.stack([zeros, vv * np.sqrt(1 - r**2) * vol], -1)
    return np.stack([col1, col2], -1)

  def drift_fn(t, x):
    var = x[..., 1]
    kappa = _param(mean_reversion, t, dtype)
    th = _param(theta, t, dtype)
    return np.stack([-var / 2, kappa * (th - var)], -1)
  return drift_fn, vol_fn


def complex_step_forward_rate(discount_rate_fn):
  """f(0,t) = d/dt [r(t) t] and its derivative for analytic `r` (numpy).

  The reference obtains both by forward-mode AD (`vector_hull_white.py:209-225,
  298-300`); for the oracle a complex step (exact to rounding for analytic
  functions) gives f, and a central difference of f gives f'.
  """
  h = 1e-30

  def fwd(t):
    t = np.asarray(t, dtype=np.float64)
    z = t + 1j * h
    return np.imag(np.asarray(discount_rate_fn(z)) * z) / h

  def fwd_grad(t, eps=1e-5):
    t = np.asarray(t, dtype=np.float64)
    return (fwd(t + eps) - fwd(t - eps)) / (2 * eps)
  return fwd, fwd_grad


def hull_white_1f_closures(mean_reversion, volatility, forward_rate_fn,
                           forward_rate_grad_fn, dtype):
  """1-factor Hull-White short-rate closures (`vector_hull_white.py:275-306`).

  drift = f'(0,t) + k f(0,t) + s^2/(2k) (1 - exp(-2 k t)) - k x ; S = [[s]].
  """
  dtype = np.dtype(dtype)

  def vol_fn(t, x):
    s = _param(volatility, t, dtype)
    return s * np.ones(x.shape[:-1] + (1, 1), dtype=dtype)

  def drift_fn(t, x):
    k = _param(mean_reversion, t, dtype)
    s = _param(volatility, t, dtype)
    f = dtype.type(forward_rate_fn(t))
    fg = dtype.type(forward_rate_grad_fn(t))
    drift = fg + k * f
    drift = drift + (s**2 / 2 / k * (1 - np.exp(-2 * k * t)) - k * x)
    return drift
  return drift_fn, vol_fn


def gbm_exact_sample_paths(mean, volatility, times, initial_state=None,
                           num_samples=1, random_type=None, seed=None, skip=0,
                           dtype=np.float64, normal_draws=None):
  """`GeometricBrownianMotion.sample_paths` (exact log-normal sampler,
  `univariate_geometric_brownian_motion.py:155-317`) -> batch_shape + [N, k, 1].

  Batches as in the reference: `mean` / `volatility` of shape `batch_shape + [1]` (or batched
  `PiecewiseConstantFunc`s), `times` of shape `[k]` or `batch_shape + [k]`, `initial_state`
  broadcastable to `batch_shape + [1]`.  The normal draws carry NO batch shape (`:277-282`): every
  element of the batch is driven by the same `[k, N]` normals."""
  from oracle import draws as draws_lib
  dtype = np.dtype(dtype)
  times = np.asarray(times, dtype=dtype)
  k = times.shape[-1]
  x0 = np.ones(1, dtype) if initial_state is None else np.asarray(initial_state, dtype)
  if normal_draws is None:
    z = draws_lib.generate_mc_normal_draws(
        1, k, num_samples, draws_lib.RandomType.PSEUDO if random_type is None else random_type,
        seed=seed, dtype=dtype, skip=skip)                    # [k, N, 1]
  else:
    z = np.transpose(np.asarray(normal_draws, dtype), [1, 0, 2])
  t = np.concatenate([np.zeros(times.shape[:-1] + (1,), dtype), times], -1)

  def integ(p, square=False):
    if callable(p):
      q = p if not square else PiecewiseConstantFunc(p.jump_locations(), p.values()**2, dtype=dtype)
      return q.integrate(t[..., :-1], t[..., 1:])
    v = np.asarray(p, dtype)
    return (v * v if square else v) * (t[..., 1:] - t[..., :-1])
  mean_int = np.expand_dims(integ(mean), -2)                  # batch_shape + [1, k]
  vol2_int = np.expand_dims(integ(volatility, square=True), -2)
  with np.errstate(invalid='ignore'):
    root = np.where(vol2_int > 0, np.sqrt(np.maximum(vol2_int, 0)), 0)      # _sqrt_no_nan
  log_inc = (mean_int - vol2_int / 2) + root * z[:, :, 0].T   # batch_shape + [N, k]
  lower = np.tril(np.ones((k, k), dtype))
  cumsum = log_inc @ lower.T
  return (np.expand_dims(x0, -1) * np.exp(cumsum))[..., None].astype(dtype)


def mvgbm_exact_sample_paths(means, volatilities, corr_matrix, times,
                             initial_state=None, num_samples=1, random_type=None,
                             seed=None, skip=0, dtype=np.float64):
  """`MultivariateGeometricBrownianMotion.sample_paths`
  (`multivariate_geometric_brownian_motion.py:153-282`) -> [N, k, dim]."""
  from oracle import draws as draws_lib
  dtype = np.dtype(dtype)
  means = np.asarray(means, dtype)
  vols = np.asarray(volatilities, dtype)
  dim = means.shape[0]
  times = np.asarray(times, dtype=dtype)
  k = times.shape[0]
  x0 = np.ones(dim, dtype) if initial_state is None else np.asarray(initial_state, dtype)
  z = draws_lib.generate_mc_normal_draws(
      dim, k, num_samples, draws_lib.RandomType.PSEUDO if random_type is None else random_type,
      seed=seed, dtype=dtype, skip=skip)                      # [k, N, dim]
  t = np.concatenate([np.zeros(1, dtype), times])
  dt = (t[1:] - t[:-1])[:, None, None]
  if corr_matrix is not None:
    chol = np.linalg.cholesky(np.asarray(corr_matrix, dtype)).astype(dtype)
    z = np.einsum('ij,knj->kni', chol, z).astype(dtype)
  log_inc = (means - vols**2 / 2) * dt + np.sqrt(dt) * vols * z    # [k, N, dim]
  cumsum = np.cumsum(log_inc, axis=0)
  return (x0 * np.exp(np.transpose(cumsum, [1, 0, 2]))).astype(dtype)
