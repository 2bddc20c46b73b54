"""Correlated multi-factor Hull-White model
(`tf_quant_finance/models/hull_white/vector_hull_white.py:35-1088`).

  dr_i = (theta_i(t) - a_i r_i) dt + sigma_i(t) dW_i,   dW_i dW_j = rho_ij dt

* `dim == 1` is `HullWhiteModel1F` (fused TQF_MODEL_HW1F kernel);
* `2 <= dim <= 4` with constant mean reversions, constant or piecewise-constant
  volatilities and a constant or piecewise-constant correlation matrix: the
  exact OU discretisation of the reference (`_sample_paths` 641-781).  One
  step is  x' = e^{-a dt} x + E[x] + sqrt(Var x) (L z),  L = cholesky(rho): an
  affine map with state-independent noise, which is what TQF_MODEL_AFFINE_ND
  steps (`x' = (x + dt (a0 + A1 x)) + B (sqrt_dt z)` with dt = sqrt_dt = 1,
  a0 = E[x], A1 = diag(e^{-a dt} - 1), B = diag(sqrt Var x) L);
* any generic callable parameter: the Euler scheme on the model's drift and
  volatility closures (`vector_hull_white.py:275-306, 406-433`).
"""
import numpy as np
import torch

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import engine
from tff_b200.math import piecewise
from tff_b200.models import euler_sampling
from tff_b200.models.hull_white import _exact
from tff_b200.models.hull_white import one_factor


def _per_factor(param, dim, dtype, name):
  """Splits a `[dim]` constant / batched piecewise-constant parameter into one
  batch-free `PiecewiseConstantFunc` per factor (`_input_type` 1033-1088).
  Returns (list or None, is_generic)."""
  if getattr(param, 'is_piecewise_constant', False):
    jumps = np.asarray(param.jump_locations(), dtype=dtype)
    values = np.asarray(param.values(), dtype=dtype)
    if jumps.ndim > 2:
      raise ValueError(
          'Batch rank of `jump_locations` should be `1` for all piecewise '
          'constant arguments but {} instead'.format(jumps.ndim - 1))
    if jumps.ndim == 2:
      if jumps.shape[0] != dim:
        raise ValueError(
            'Batch shape of `jump_locations` should be either empty or '
            '`[{0}]` but `[{1}]` instead'.format(dim, jumps.shape[0]))
      return [piecewise.PiecewiseConstantFunc(jumps[i], values[i], dtype=dtype)
              for i in range(dim)], False
    if dim != 1:
      raise ValueError(
          'Batch shape of `jump_locations` should be `[{0}]` for a {0}-factor '
          'model but is empty'.format(dim))
    return [param], False
  if callable(param):
    return None, True
  value = np.asarray(_tensor.to_numpy(param, dtype), dtype=dtype).reshape(-1)
  if value.shape[0] != dim:
    raise ValueError('Length of {} ({}) should be the same as `dims`({}).'.format(
        name, value.shape[0], dim))
  return [piecewise.PiecewiseConstantFunc([], [value[i]], dtype=dtype)
          for i in range(dim)], False


class _VectorExactSpec(engine.ModelSpec):
  """Per-step table of TQF_MODEL_AFFINE_ND carrying the exact OU step."""

  def __init__(self, tables, corr_root_fn):
    d = len(tables)
    self.kind, self.dim, self.num_factors = _lib.MODEL_AFFINE_ND, d, d
    self.num_coef = 2 + d + 2 * d * d
    self.tables, self.corr_root_fn = tables, corr_root_fn

  def coef_table(self, all_times, dtype):
    d = self.dim
    t = np.asarray(all_times, dtype=dtype)
    dt = t[1:] - t[:-1]
    s = dt.shape[0]
    rows = np.zeros((s, self.num_coef), dtype=np.float64)
    rows[:, 0] = 1.0
    rows[:, 1] = 1.0
    root = self.corr_root_fn(t, dt)                       # [S + 1, d, d] or None
    for i, tab in enumerate(self.tables):
      rows[:, 2 + i] = tab.conditional_mean_x(t)
      rows[:, 2 + d + i * d + i] = np.exp(-tab.k * dt) - 1.0
      var = tab.conditional_variance_x(t)
      c = np.sqrt(np.maximum(var, 0))
      c = np.where(c > 0.0, c, 0.0)
      for j in range(d):
        lij = (1.0 if i == j else 0.0) if root is None else root[:s, i, j]
        rows[:, 2 + d + d * d + i * d + j] = c * lij
    return rows


class VectorHullWhiteModel:
  """Ensemble of correlated Hull-White short-rate models."""

  def __init__(self, dim, mean_reversion, volatility, initial_discount_rate_fn,
               corr_matrix=None, dtype=None, name=None):
    self._name = name or 'hull_white_model'
    self._dim = int(dim)
    self._dtype = _tensor.np_dtype(dtype, np.float32)
    dt_ = self._dtype
    self._initial_discount_rate_fn = initial_discount_rate_fn
    self._one_factor = None
    if self._dim == 1:
      if corr_matrix is not None and not callable(corr_matrix):
        corr_matrix = None                # a 1 x 1 correlation matrix is [[1]]
      self._one_factor = one_factor.HullWhiteModel1F(
          mean_reversion, volatility, initial_discount_rate_fn, dtype=dt_, name=name)
      return

    def rate_i(i):
      def fn(t):
        r = initial_discount_rate_fn(t)
        nd = t.dim() if isinstance(t, torch.Tensor) else np.ndim(t)
        rd = r.dim() if isinstance(r, torch.Tensor) else np.ndim(r)
        return r[..., i] if rd == nd + 1 else r
      return fn
    self._rate_fns = [rate_i(i) for i in range(self._dim)]
    fns = [_exact.forward_rate_fns(f, dt_) for f in self._rate_fns]
    self._fwd = [f[0] for f in fns]
    self._fwd_grad = [f[1] for f in fns]

    self._mean_reversion, self._volatility = mean_reversion, volatility
    mr, g1 = _per_factor(mean_reversion, self._dim, dt_, 'mean_reversion')
    vol, g2 = _per_factor(volatility, self._dim, dt_, 'volatility')
    mr_jumps = (not g1) and any(np.asarray(m.jump_locations()).size for m in mr)
    self._mr_fns, self._vol_fns = mr, vol
    # correlation: None | constant [dim, dim] | piecewise (rank-1 jumps) | generic callable
    self._corr = corr_matrix
    g3 = False
    if corr_matrix is not None:
      if getattr(corr_matrix, 'is_piecewise_constant', False):
        if np.asarray(corr_matrix.jump_locations()).ndim != 1:
          raise ValueError('Batch rank of `jump_locations` should be `0` for '
                           'the correlation matrix.')
      elif callable(corr_matrix):
        g3 = True
      else:
        c = _tensor.to_numpy(corr_matrix, dt_)
        if c.shape != (self._dim, self._dim):
          raise ValueError('`corr_matrix` should have shape [{0}, {0}] but is {1}'.format(
              self._dim, list(c.shape)))
        self._corr = c
    self._sample_with_generic = g1 or g2 or g3 or mr_jumps
    self._is_piecewise_constant = not (g1 or g2 or g3)
    self._tables = None
    if not self._sample_with_generic:
      if self._dim > 4:
        raise NotImplementedError(
            'The fused exact Hull-White sampler covers dim <= 4.')
      self._tables = [
          _exact.ExactTables(
              np.asarray(mr[i](np.zeros(1, dt_))).reshape(-1)[0], vol[i], dt_)
          for i in range(self._dim)]

  # ------------------------------------------------------------ accessors --
  def dim(self):
    return self._dim

  def dtype(self):
    return self._dtype

  def name(self):
    return self._name

  @property
  def mean_reversion(self):
    return self._one_factor.mean_reversion if self._one_factor else self._mean_reversion

  @property
  def volatility(self):
    return self._one_factor.volatility if self._one_factor else self._volatility

  def instant_forward_rate(self, t):
    """f(0, t) of shape `t.shape + [dim]` (numpy)."""
    if self._one_factor:
      return np.asarray(self._one_factor.instant_forward_rate(t))[..., None]
    t = np.asarray(t, dtype=self._dtype)
    return np.stack([f(t) for f in self._fwd], -1)

  # ------------------------------------------------------------ helpers ----
  def _corr_at(self, t):
    """Correlation matrices `[n, dim, dim]` at times `t` (numpy)."""
    d = self._dim
    if self._corr is None:
      return None
    if getattr(self._corr, 'is_piecewise_constant', False):
      return np.asarray(self._corr(t), dtype=self._dtype).reshape(t.shape[0], d, d)
    if callable(self._corr):
      out = []
      for ti in t:
        try:
          c = self._corr(torch.tensor(float(ti), dtype=_tensor.torch_dtype(self._dtype)))
          c = c.detach().cpu().numpy() if isinstance(c, torch.Tensor) else c
        except Exception:  # pylint: disable=broad-except
          c = self._corr(self._dtype.type(ti))
        out.append(np.asarray(c, dtype=self._dtype))
      return np.stack(out, 0)
    return np.broadcast_to(self._corr, (t.shape[0], d, d))

  def _corr_root(self, t, dt):
    """cholesky(rho(t + min(dt) / 2)) (`vector_hull_white.py:700-703`)."""
    if self._corr is None or dt.shape[0] == 0:
      return None
    return np.linalg.cholesky(self._corr_at(t + dt.min() / 2).astype(np.float64))

  def _prepare_grid(self, times, times_grid):
    """`_prepare_grid` (`vector_hull_white.py:982-1030`): the jump locations of
    every factor's parameters join the grid, duplicates included."""
    dt_ = self._dtype
    if times_grid is None:
      jumps = [np.asarray(p.jump_locations(), dtype=dt_).reshape(-1)
               for p in list(self._mr_fns) + list(self._vol_fns)]
      # parameter order of the reference: all mean-reversion jumps, then all
      # volatility jumps (the sort makes the order immaterial)
      all_times = np.sort(np.concatenate([np.zeros(1, dt_), times] + jumps),
                          kind='stable').astype(dt_)
      idx = np.searchsorted(all_times, times, side='left')
    else:
      all_times = np.asarray(times_grid, dtype=dt_)
      idx = np.minimum(np.searchsorted(all_times, times, side='left'),
                       all_times.shape[0] - 1)
      d1 = all_times[idx] - times
      d2 = all_times[np.maximum(idx - 1, 0)] - times
      idx = np.where(np.abs(d2) > np.abs(d1), idx, np.maximum(idx - 1, 0))
    mask = np.zeros(all_times.shape[0], dtype=bool)
    mask[idx] = True
    return all_times, mask, idx

  def _closures(self):
    """numpy (drift_fn, volatility_fn) (`vector_hull_white.py:275-306`)."""
    d, dt_ = self._dim, self._dtype

    def params(fns, generic, t):
      if generic is not None:
        try:
          v = generic(torch.tensor(float(t), dtype=_tensor.torch_dtype(dt_)))
          v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v
        except Exception:  # pylint: disable=broad-except
          v = generic(dt_.type(t))
        return np.asarray(v, dtype=np.float64).reshape(d)
      return np.asarray([np.asarray(f(np.asarray([t], dt_))).reshape(-1)[0] for f in fns],
                        dtype=np.float64)
    mr_generic = self._mean_reversion if self._mr_fns is None else None
    vol_generic = self._volatility if self._vol_fns is None else None

    def drift_fn(t, x):
      t = float(t)
      x = np.asarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x,
                     dtype=np.float64)
      k = params(self._mr_fns, mr_generic, t)
      s = params(self._vol_fns, vol_generic, t)
      f = np.asarray([fn(np.asarray(t)) for fn in self._fwd], dtype=np.float64).reshape(d)
      fg = np.asarray([fn(np.asarray(t)) for fn in self._fwd_grad], dtype=np.float64).reshape(d)
      return fg + k * f + s**2 / 2 / k * (1 - np.exp(-2 * k * t)) - k * x

    def vol_fn(t, x):
      t = float(t)
      s = params(self._vol_fns, vol_generic, t)
      if self._corr is None:
        m = np.diag(s)
      else:
        root = np.linalg.cholesky(self._corr_at(np.asarray([t], dt_))[0].astype(np.float64))
        m = s[:, None] * root
      n = np.shape(x)[0] if np.ndim(x) > 1 else 1
      return np.broadcast_to(m, (n, d, d))
    return drift_fn, vol_fn

  # ------------------------------------------------------------ sampling ---
  def sample_paths(self, times, num_samples=1, random_type=None, seed=None,
                   skip=0, time_step=None, times_grid=None, normal_draws=None,
                   validate_args=False, name=None):
    """Short-rate paths `[num_samples, k, dim]` (`vector_hull_white.py:319-449`)."""
    if self._one_factor:
      return self._one_factor.sample_paths(
          times, num_samples, random_type, seed, skip, time_step, times_grid,
          normal_draws, validate_args, name)
    del name, validate_args
    dt_ = self._dtype
    times = _tensor.to_numpy(times, dt_)
    if times.ndim != 1:
      raise ValueError('`times` should be a rank 1 Tensor. '
                       'Rank is {} instead.'.format(times.ndim))
    if self._sample_with_generic:
      if time_step is None and times_grid is None:
        raise ValueError(
            'Either `time_step` or `times_grid` has to be specified when '
            'at least one of the parameters is a generic callable.')
      drift_fn, vol_fn = self._closures()
      x0 = self.instant_forward_rate(np.zeros((), dt_)).reshape(self._dim)
      return euler_sampling.sample(
          self._dim, drift_fn, vol_fn, times, time_step=time_step,
          num_samples=num_samples, initial_state=x0, random_type=random_type,
          seed=seed, skip=skip, times_grid=times_grid, normal_draws=normal_draws,
          dtype=dt_)
    if normal_draws is not None:
      normal_draws = _tensor.from_dlpack(normal_draws)
      num_samples = int(normal_draws.shape[0])
      if int(normal_draws.shape[2]) != self._dim:
        raise ValueError(
            '`dim` should be equal to `normal_draws.shape[2]` but are '
            '{0} and {1} respectively'.format(self._dim, int(normal_draws.shape[2])))
    grid = None if times_grid is None else _tensor.to_numpy(times_grid, dt_)
    all_times, mask, _ = self._prepare_grid(times, grid)
    if normal_draws is not None and int(normal_draws.shape[1]) != all_times.shape[0] - 1:
      raise ValueError(
          '`tf.shape(normal_draws)[1]` should be equal to the number of all '
          '`times` plus the number of all jumps of the piecewise constant '
          'parameters.')
    num_steps, record_slot = engine.record_plan(mask, times.shape[0])
    spec = _VectorExactSpec(self._tables, self._corr_root)
    rng = engine.RngSpec(random_type, seed, skip, normal_draws)
    plan = engine.Plan(spec, all_times, num_steps, np.zeros(self._dim, dt_), rng,
                       int(num_samples), dt_)
    try:
      x = plan.paths(record_slot, times.shape[0])             # [N, k, dim]
    finally:
      plan.close()
    f0 = torch.as_tensor(self.instant_forward_rate(times), device=x.device, dtype=x.dtype)
    return x + f0[None, :, :]

  def sample_discount_curve_paths(self, times, curve_times, num_samples=1, random_type=None,
                                  seed=None, skip=0, time_step=None, times_grid=None,
                                  normal_draws=None, validate_args=False, name=None):
    """Simulated discount curves (`vector_hull_white.py:451-592`):
    `(P(t, t + tau) [num_samples, m, k, dim], short rates [num_samples, k, dim])`,
    factor d discounting on its own curve.  Needs piecewise-constant (or constant)
    parameters, as the reference."""
    if self._one_factor:
      return self._one_factor.sample_discount_curve_paths(
          times, curve_times, num_samples, random_type, seed, skip, time_step, times_grid,
          normal_draws, validate_args, name)
    del name
    if not self._is_piecewise_constant or self._sample_with_generic:
      raise ValueError('All paramaters `mean_reversion`, `volatility`, and '
                       '`corr_matrix`must be piecewise constant functions.')
    dt_ = self._dtype
    times = _tensor.to_numpy(times, dt_)
    curve_times = _tensor.to_numpy(curve_times, dt_)
    rates = self.sample_paths(times, num_samples, random_type, seed, skip, time_step,
                              times_grid, normal_draws, validate_args)
    rate_fns = [(lambda t, f=f: _exact.discount_rate(f, t, dt_)) for f in self._rate_fns]
    curves = one_factor.discount_curves_on_device(
        rates, times, curve_times, [tab.k for tab in self._tables],
        [tab.y_t(times) for tab in self._tables], rate_fns, self._fwd, dt_)
    return curves, rates
