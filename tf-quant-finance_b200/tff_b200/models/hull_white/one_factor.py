"""One-factor Hull-White model (`models/hull_white/one_factor.py`,
`models/hull_white/vector_hull_white.py` with dim = 1).

  dr = (theta(t) - a r) dt + sigma(t) dW

* constant mean reversion and constant / piecewise-constant volatility: the
  exact OU discretisation of the reference (`_sample_paths` 641-781) runs in
  the fused kernel (TQF_MODEL_HW1F), which also carries the path integral of
  the short rate used by the swaption pricer;
* any generic callable parameter (or a mean reversion with jumps): the Euler
  scheme through `euler_sampling.sample` with `initial_state = f(0, 0)`
  (`vector_hull_white.py:406-433`), as the reference does.
"""
import numpy as np
import torch

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import engine
from tff_b200.math import piecewise
from tff_b200.models import closures
from tff_b200.models import euler_sampling
from tff_b200.models import generic_ito_process
from tff_b200.models.hull_white import _exact


def _input_type(param, dtype, name):
  """`_input_type` (`vector_hull_white.py:1033-1088`) for dim = 1."""
  if getattr(param, 'is_piecewise_constant', False):
    jumps = np.asarray(param.jump_locations())
    if jumps.ndim > 2:
      raise ValueError(
          'Batch rank of `jump_locations` should be `1` for all piecewise '
          'constant arguments but {} instead'.format(jumps.ndim - 1))
    generic = name == 'mean_reversion' and jumps.reshape(-1).shape[0] > 0
    return param, generic, True
  if callable(param):
    return param, True, False
  value = np.asarray(_tensor.to_numpy(param, dtype)).reshape(-1)
  if value.shape[0] != 1:
    raise ValueError('Length of {} ({}) should be the same as `dims`({}).'.format(
        name, value.shape[0], 1))
  return piecewise.PiecewiseConstantFunc([], value, dtype=dtype), False, True


def discount_curves_on_device(rates, times, curve_times, mean_reversions, y_t, rate_fns, fwd_fns,
                              dtype):
  """P(t_j, t_j + tau_i) `[N, m, k, dim]` along the short-rate paths `rates`
  (`[N, k, dim]` device tensor, any strides): the path-independent factors of
  `_bond_reconstitution` (`vector_hull_white.py:783-814`) are tabulated on the host
  per factor d,
    G[i, j, d] = (1 - e^{-a_d tau_i}) / a_d,
    A[i, j, d] = P0_d(t_j + tau_i) / P0_d(t_j) exp(-y_d(t_j) G^2 / 2),
  and `tqf_hw_discount_curves` writes A exp(-(r - f0) G) in one pass."""
  import ctypes as C  # pylint: disable=g-import-not-at-top
  k, m, dim = times.shape[0], curve_times.shape[0], len(mean_reversions)
  t = times[None, :]
  big_t = t + curve_times[:, None]                                       # [m, k]
  a_tab = np.empty((m, k, dim), dtype=np.float64)
  g_tab = np.empty((m, k, dim), dtype=np.float64)
  f0 = np.empty((k, dim), dtype=np.float64)
  for d in range(dim):
    kap = dtype.type(mean_reversions[d])
    g = ((1. - np.exp(-kap * (big_t - t))) / kap).astype(dtype)
    p0t = np.exp(-rate_fns[d](times) * times).astype(dtype)
    p0T = np.exp(-rate_fns[d](big_t) * big_t).astype(dtype)
    y = np.asarray(y_t[d], dtype=dtype)[None, :]
    # the reference multiplies p_0_t_tau by exp(-term1 - 0.5 term2); exp(a + b) is
    # split here into exp(a) exp(b): one rounding apart
    a_tab[:, :, d] = (p0T / p0t[None, :]) * np.exp(-0.5 * (y * g**2))
    g_tab[:, :, d] = g
    f0[:, d] = np.asarray(fwd_fns[d](times), dtype=dtype)
  dev = rates.device
  f0_d = torch.as_tensor(f0, device=dev)
  a_d = torch.as_tensor(np.ascontiguousarray(a_tab), device=dev)
  g_d = torch.as_tensor(np.ascontiguousarray(g_tab), device=dev)
  n = int(rates.shape[0])
  out = torch.empty((n, m, k, dim), dtype=rates.dtype, device=dev)
  _lib.require_cuda()
  sp, st, sd = rates.stride()
  _lib.check(_lib.lib().tqf_hw_discount_curves(
      rates.data_ptr(), sp, st, sd, f0_d.data_ptr(), a_d.data_ptr(), g_d.data_ptr(), n, m, k,
      dim, _tensor.tqf_dtype(dtype), out.data_ptr(), _tensor.current_stream_ptr()))
  return out


class HullWhite1FSpec(engine.ModelSpec):
  """Per-step table of TQF_MODEL_HW1F: A, B, C, W, W f(0, t_{i+1}); state [x, I]."""
  kind, dim, num_factors, num_coef = _lib.MODEL_HW1F, 2, 1, 5

  def __init__(self, tables, fwd, integral_weights=None):
    self.tables, self.fwd = tables, fwd
    self.integral_weights = integral_weights       # per step, or None -> 0

  def coef_table(self, all_times, dtype):
    t = np.asarray(all_times, dtype=dtype)
    dt = t[1:] - t[:-1]
    a = np.exp(-self.tables.k * dt)
    b = self.tables.conditional_mean_x(t)
    var = self.tables.conditional_variance_x(t)
    c = np.sqrt(np.maximum(var, 0))
    c = np.where(c > 0.0, c, 0.0)
    w = (np.zeros_like(dt) if self.integral_weights is None
         else np.asarray(self.integral_weights, dtype=dtype))
    wf = w * self.fwd(t[1:])
    return np.stack([a, b, c, w, wf], -1).astype(np.float64)


class HullWhiteModel1F(generic_ito_process.GenericItoProcess):
  """One-factor Hull-White short-rate model."""

  def __init__(self, mean_reversion, volatility, initial_discount_rate_fn,
               dtype=None, name=None):
    self._name = name or 'hull_white_one_factor'
    self._dtype = _tensor.np_dtype(dtype, np.float32)
    self._dim = 1
    self._initial_discount_rate_fn = initial_discount_rate_fn
    self._fwd, self._fwd_grad = _exact.forward_rate_fns(
        initial_discount_rate_fn, self._dtype)
    self._mean_reversion, g1, p1 = _input_type(mean_reversion, self._dtype,
                                               'mean_reversion')
    self._volatility, g2, p2 = _input_type(volatility, self._dtype, 'volatility')
    self._sample_with_generic = g1 or g2
    self._is_piecewise_constant = p1 and p2
    self._tables = None
    if self._is_piecewise_constant and not self._sample_with_generic:
      k = np.asarray(self._mean_reversion(np.zeros(1, self._dtype))).reshape(-1)[0]
      self._tables = _exact.ExactTables(k, self._volatility, self._dtype)

    mr, vol, fwd, fwd_grad, dt_ = (self._mean_reversion, self._volatility,
                                   self._fwd, self._fwd_grad, self._dtype)

    def _scalar(p, t):
      return np.asarray(p(np.asarray(t, dtype=dt_).reshape(-1)), dtype=dt_).reshape(np.shape(t))

    def a0(t):   # f'(0,t) + k f(0,t) + s^2/(2k) (1 - exp(-2 k t))   (lines 292-306)
      k, s = _scalar(mr, t), _scalar(vol, t)
      return fwd_grad(t) + k * fwd(t) + s**2 / 2 / k * (1 - np.exp(-2 * k * t))

    def a1(t):
      return -_scalar(mr, t)

    def b(t):
      return _scalar(vol, t)
    drift_fn, vol_fn = closures.affine_closures(a0, a1, b)
    super().__init__(1, drift_fn, vol_fn, self._dtype, self._name)

  @property
  def mean_reversion(self):
    return self._mean_reversion

  @property
  def volatility(self):
    return self._volatility

  def instant_forward_rate(self, t):
    """f(0, t) as a numpy array."""
    return self._fwd(np.asarray(t, dtype=self._dtype))

  # ------------------------------------------------------------ grids -----
  def _prepare_grid(self, times, times_grid):
    """`_prepare_grid` (`vector_hull_white.py:982-1030`)."""
    dt_ = self._dtype
    if times_grid is None:
      jumps = [np.asarray(self._volatility.jump_locations(), dtype=dt_).reshape(-1),
               np.asarray(self._mean_reversion.jump_locations(), dtype=dt_).reshape(-1)]
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

  def _exact_plan(self, times, num_samples, random_type, seed, skip, times_grid,
                  normal_draws, integral_weights_fn=None):
    all_times, mask, idx = self._prepare_grid(times, times_grid)
    num_steps, record_slot = engine.record_plan(mask, times.shape[0])
    weights = None if integral_weights_fn is None else integral_weights_fn(all_times, idx)
    spec = HullWhite1FSpec(self._tables, self._fwd, weights)
    rng = engine.RngSpec(random_type, seed, skip, normal_draws)
    plan = engine.Plan(spec, all_times, num_steps, np.zeros(2, self._dtype), rng,
                       num_samples, self._dtype)
    return plan, record_slot, all_times, idx

  # --------------------------------------------------------- sampling -----
  def sample_paths(self, times, num_samples=1, random_type=None, seed=None,
                   skip=0, time_step=None, times_grid=None, normal_draws=None,
                   validate_args=False, name=None):
    """Short-rate paths `[num_samples, k, 1]` (`vector_hull_white.py:319-449`)."""
    del name
    times = _tensor.to_numpy(times, self._dtype)
    if times.ndim != 1:
      raise ValueError('`times` should be a rank 1 Tensor. '
                       'Rank is {} instead.'.format(times.ndim))
    if self._sample_with_generic:
      if time_step is None and times_grid is None:
        raise ValueError(
            'Either `time_step` or `times_grid` has to be specified when '
            'at least one of the parameters is a generic callable.')
      x0 = self._fwd(np.zeros(1, self._dtype))
      return euler_sampling.sample(
          1, self._drift_fn, self._volatility_fn, times, time_step=time_step,
          num_samples=num_samples, initial_state=x0, random_type=random_type,
          seed=seed, skip=skip, times_grid=times_grid, normal_draws=normal_draws,
          dtype=self._dtype)
    if normal_draws is not None:
      normal_draws = _tensor.from_dlpack(normal_draws)
      num_samples = int(normal_draws.shape[0])
      if int(normal_draws.shape[2]) != 1:
        raise ValueError(
            '`dim` should be equal to `normal_draws.shape[2]` but are '
            '{0} and {1} respectively'.format(1, int(normal_draws.shape[2])))
    plan, record_slot, _, _ = self._exact_plan(
        times, int(num_samples), random_type, seed, skip,
        None if times_grid is None else _tensor.to_numpy(times_grid, self._dtype),
        normal_draws)
    del validate_args
    try:
      state = plan.paths(record_slot, times.shape[0])       # [N, k, 2] = (x, I)
    finally:
      plan.close()
    f0 = torch.as_tensor(self._fwd(times), device=state.device, dtype=state.dtype)
    return state[..., 0:1] + f0[None, :, None]

  def _y_and_k(self, times):
    return self._tables.y_t(times), self._tables.k

  def _bond_reconstitution(self, times, maturities, short_rate, y_t):
    """`_bond_reconstitution` (`vector_hull_white.py:783-814`) with torch ops on
    device tensors (`short_rate` broadcastable to `times`)."""
    dev, td = short_rate.device, short_rate.dtype
    k = float(self._tables.k)

    def dv(a):
      return torch.as_tensor(np.array(a, dtype=self._dtype), device=dev, dtype=td)
    f0 = dv(self._fwd(times))
    p0t = dv(np.exp(-_exact.discount_rate(self._initial_discount_rate_fn, times, self._dtype) * times))
    p0T = dv(np.exp(-_exact.discount_rate(self._initial_discount_rate_fn, maturities, self._dtype) * maturities))
    g = dv((1. - np.exp(-k * (maturities - times))) / k)
    x_t = short_rate - f0
    return (p0T / p0t) * torch.exp(-x_t * g - 0.5 * dv(y_t) * g**2)

  def sample_discount_curve_paths(self, times, curve_times, num_samples=1,
                                  random_type=None, seed=None, skip=0,
                                  time_step=None, times_grid=None,
                                  normal_draws=None, validate_args=False,
                                  name=None):
    """(P(t, t + tau) `[N, m, k, 1]`, short rates `[N, k, 1]`)
    (`vector_hull_white.py:451-592`)."""
    del name
    if not self._is_piecewise_constant or self._sample_with_generic:
      raise ValueError('All paramaters `mean_reversion`, `volatility`, and '
                       '`corr_matrix`must be piecewise constant functions.')
    times = _tensor.to_numpy(times, self._dtype)
    curve_times = _tensor.to_numpy(curve_times, self._dtype)
    rates = self.sample_paths(times, num_samples, random_type, seed, skip,
                              time_step, times_grid, normal_draws, validate_args)
    rate = lambda t: _exact.discount_rate(self._initial_discount_rate_fn, t, self._dtype)
    curves = discount_curves_on_device(
        rates, times, curve_times, [self._tables.k], [self._tables.y_t(times)], [rate],
        [self._fwd], self._dtype)
    return curves, rates

  def discount_bond_price(self, short_rate, times, maturities, name=None):
    """P(t, T) given r(t) (`vector_hull_white.py:594-636`); numpy in, numpy out."""
    del name
    times = _tensor.to_numpy(times, self._dtype)
    maturities = _tensor.to_numpy(maturities, self._dtype)
    r = torch.as_tensor(_tensor.to_numpy(short_rate, self._dtype))[..., 0]
    y_t = self._tables.y_t(times.reshape(-1)).reshape(times.shape)
    return self._bond_reconstitution(times, maturities, r, y_t)[..., None].numpy()
