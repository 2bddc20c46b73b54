"""Oracle (test infrastructure): one-factor Hull-White model.

Restates `models/hull_white/vector_hull_white.py` for dim = 1:
  * exact discretisation tables `_conditional_mean_x` 876-907,
    `_conditional_variance_x` 925-949, `_compute_yt` 857-874, the integrals
    909-923/951-955 and `_exact_discretization_setup` 816-855;
  * `_prepare_grid` 982-1030 and the sampling loop `_sample_paths` 641-781;
  * `_bond_reconstitution` 783-814, `discount_bond_price` 594-636,
    `sample_discount_curve_paths` 451-592;
and `models/hjm/swaption_util.py:28-170` +
`models/hull_white/swaption.py:216-312` (Monte-Carlo swaption price).

`initial_discount_rate_fn` must be a numpy function analytic in `t`; the
instantaneous forward rate f(0,t) = d/dt [r(t) t], obtained by forward-mode AD
in the reference (lines 209-225), is computed here with a complex step.
"""
import numpy as np

from oracle import draws as draws_lib
from oracle import models as models_lib


class HullWhiteModel1F:
  """Constant mean reversion, constant or piecewise-constant volatility."""

  def __init__(self, mean_reversion, volatility, initial_discount_rate_fn,
               dtype=np.float64):
    self.dtype = np.dtype(dtype)
    self.k = self.dtype.type(mean_reversion)
    if isinstance(volatility, models_lib.PiecewiseConstantFunc):
      self.vol = volatility
    else:
      self.vol = models_lib.PiecewiseConstantFunc(
          [], [volatility], dtype=self.dtype)
    self.rate_fn = initial_discount_rate_fn
    self.fwd, self.fwd_grad = models_lib.complex_step_forward_rate(
        initial_discount_rate_fn)
    # _exact_discretization_setup: knots = sort(vol jumps ++ mr jumps)
    self.jumps = np.sort(np.asarray(self.vol.jump_locations(), dtype=self.dtype))
    self.jump_vol = self.vol(self.jumps)
    self.jump_mr = np.full(self.jumps.shape, self.k, dtype=self.dtype)
    self.padded_knots = np.concatenate(
        [np.zeros(1, dtype=self.dtype), self.jumps[:-1]])[:self.jumps.shape[0]]

  # --- integrals ------------------------------------------------------------
  @staticmethod
  def _y_integral(t0, t, vol, k):
    return (vol * vol) / (2 * k) * (np.exp(2 * k * t) - np.exp(2 * k * t0))

  @staticmethod
  def _ex_integral(t0, t, vol, k, y_t0):
    value = (np.exp(k * t) - np.exp(k * t0) + np.exp(2 * k * t0) *
             (np.exp(-k * t) - np.exp(-k * t0)))
    return value * vol**2 / (2 * k * k) + y_t0 * (np.exp(-k * t0) - np.exp(-k * t)) / k

  def _y_at_knots(self):
    between = self._y_integral(self.padded_knots, self.jumps, self.jump_vol,
                               self.jump_mr)
    return np.concatenate([np.zeros(1, dtype=self.dtype), np.cumsum(between)])

  def compute_yt(self, t):
    t = np.asarray(t, dtype=self.dtype)
    sigma_t, mr_t = self.vol(t), self.k
    idx = np.searchsorted(self.jumps, t, side='left')
    y_at = self._y_at_knots()
    vn = np.concatenate([np.zeros(1, dtype=self.dtype), self.jumps])
    y_t = self._y_integral(vn[idx], t, sigma_t, mr_t) + y_at[idx]
    return np.exp(-2 * mr_t * t) * y_t

  def conditional_mean_x(self, t):
    t = np.asarray(t, dtype=self.dtype)
    sigma_t, mr_t = self.vol(t), self.k
    idx = np.searchsorted(self.jumps, t, side='left')
    vn = np.concatenate([np.zeros(1, dtype=self.dtype), self.jumps])
    y_at = self._y_at_knots()
    ex_between = self._ex_integral(self.padded_knots, self.jumps, self.jump_vol,
                                   self.jump_mr, y_at[:-1])
    ex_at = np.concatenate([np.zeros(1, dtype=self.dtype), np.cumsum(ex_between)])
    ex = self._ex_integral(vn[idx], t, sigma_t, mr_t, y_at[idx]) + ex_at[idx]
    return (ex[1:] - ex[:-1]) * np.exp(-mr_t * t[1:])

  def conditional_variance_x(self, t):
    t = np.asarray(t, dtype=self.dtype)
    sigma_t, mr_t = self.vol(t), self.k
    between = self._y_integral(self.padded_knots, self.jumps, self.jump_vol,
                               self.jump_mr)
    var_at = np.concatenate([np.zeros(1, dtype=self.dtype), np.cumsum(between)])
    idx = np.searchsorted(self.jumps, t, side='left')
    vn = np.concatenate([np.zeros(1, dtype=self.dtype), self.jumps])
    var = self._y_integral(vn[idx], t, sigma_t, mr_t) + var_at[idx]
    return (var[1:] - var[:-1]) * np.exp(-2 * mr_t * t[1:])

  # --- sampling -------------------------------------------------------------
  def prepare_grid(self, times, times_grid=None):
    times = np.asarray(times, dtype=self.dtype)
    if times_grid is None:
      all_times = np.sort(np.concatenate(
          [np.zeros(1, dtype=self.dtype), times, self.jumps, np.zeros(0)]),
                          kind='stable').astype(self.dtype)
      idx = np.searchsorted(all_times, times, side='left')
    else:
      all_times = np.asarray(times_grid, dtype=self.dtype)
      idx = np.searchsorted(all_times, times, side='left')
      idx = np.minimum(idx, all_times.shape[0] - 1)
      d1 = all_times[idx] - times
      d2 = all_times[np.maximum(idx - 1, 0)] - times
      idx = np.where(np.abs(d2) > np.abs(d1), idx, np.maximum(idx - 1, 0))
    mask = np.zeros(all_times.shape[0], dtype=bool)
    mask[idx] = True
    return all_times, mask

  def sample_paths(self, times, num_samples, random_type=None, seed=None,
                   skip=0, times_grid=None, normal_draws=None, path_range=None):
    """Short-rate paths [num_samples, k, 1] (`_sample_paths` 641-781).

    PSEUDO types are given the precomputed-draws layout of the STATELESS types
    (the reference draws them step by step from a stateful op, which is not
    reproducible)."""
    times = np.asarray(times, dtype=self.dtype)
    k = times.shape[0]
    all_times, keep_mask = self.prepare_grid(times, times_grid)
    dt = all_times[1:] - all_times[:-1]
    steps = dt.shape[0]
    if normal_draws is None:
      normal_draws = draws_lib.generate_mc_normal_draws(
          1, steps, num_samples,
          draws_lib.RandomType.PSEUDO if random_type is None else random_type,
          seed=seed, dtype=self.dtype, skip=skip, path_range=path_range)
      num_samples = normal_draws.shape[1]
    else:
      normal_draws = np.transpose(np.asarray(normal_draws, self.dtype), [1, 0, 2])
    exp_x_t = self.conditional_mean_x(all_times)
    var_x_t = self.conditional_variance_x(all_times)
    x = np.zeros((num_samples, 1), dtype=self.dtype)
    record = k != 1
    slots = [None] * k
    if record:
      slots[0] = x + self.dtype.type(self.fwd(all_times[0]))
    out = x + self.dtype.type(self.fwd(all_times[0]))
    written = int(keep_mask[0])
    i = 0
    while i < steps and written < k:
      vol = np.sqrt(np.maximum(var_x_t[i], 0))
      vol = vol if vol > 0 else self.dtype.type(0)
      x = np.exp(-self.k * dt[i]) * x + exp_x_t[i] + vol * normal_draws[i]
      out = x + self.dtype.type(self.fwd(all_times[i + 1]))
      if record:
        slots[written] = out
      written += int(keep_mask[i + 1])
      i += 1
    if not record:
      return out[:, None, :]
    # TensorArray.stack() yields zeros for slots that were never written
    # (possible when `times` holds exact duplicates)
    slots = [np.zeros_like(x) if s is None else s for s in slots]
    return np.transpose(np.stack(slots, 0), [1, 0, 2])

  # --- bonds ----------------------------------------------------------------
  def bond_reconstitution(self, times, maturities, short_rate, y_t):
    """P(t, T) (`_bond_reconstitution` 783-814); dim axis dropped."""
    times = np.asarray(times, dtype=self.dtype)
    maturities = np.asarray(maturities, dtype=self.dtype)
    x_t = short_rate - self.fwd(times)
    p_0_t = np.exp(-self.rate_fn(times) * times)
    p_0_t_tau = np.exp(-self.rate_fn(maturities) * maturities) / p_0_t
    g = (1. - np.exp(-self.k * (maturities - times))) / self.k
    return p_0_t_tau * np.exp(-x_t * g - 0.5 * y_t * g**2)

  def discount_bond_price(self, short_rate, times, maturities):
    """`discount_bond_price` 594-636: short_rate `[..., 1]` -> `[..., 1]`."""
    times = np.asarray(times, dtype=self.dtype)
    y_t = self.compute_yt(times.reshape(-1)).reshape(times.shape)
    r = np.asarray(short_rate, dtype=self.dtype)[..., 0]
    return self.bond_reconstitution(times, maturities, r, y_t)[..., None]

  def sample_discount_curve_paths(self, times, curve_times, num_samples,
                                  random_type=None, seed=None, skip=0, path_range=None):
    """(P(t, t+tau) [N, m, k, 1], r_t [N, k, 1]) (`...py:451-592`)."""
    times = np.asarray(times, dtype=self.dtype)
    curve_times = np.asarray(curve_times, dtype=self.dtype)
    y_t = self.compute_yt(times)
    rates = self.sample_paths(times, num_samples, random_type, seed, skip,
                              path_range=path_range)
    r = rates[:, None, :, 0]                                 # [N, 1, k]
    t = times[None, None, :]
    tau = curve_times[None, :, None]
    p = self.bond_reconstitution(t, t + tau, r, y_t[None, None, :])
    return p[..., None], rates


def swaption_price_mc(*, expiries, fixed_leg_payment_times,
                      fixed_leg_daycount_fractions, fixed_leg_coupon,
                      reference_rate_fn, mean_reversion, volatility,
                      notional=1.0, is_payer_swaption=True, num_samples=100,
                      random_type=None, seed=None, skip=0, time_step=None,
                      dtype=np.float64, return_payoffs=False, path_range=None):
  """`swaption_price(use_analytic_pricing=False)` (`swaption.py:216-312`) with
  `discount_factors_and_bond_prices_from_samples` (`hjm/swaption_util.py:28-170`).

  `expiries` of shape `batch`; the leg arrays of shape `batch + [m]`."""
  dtype = np.dtype(dtype)
  expiries = np.asarray(expiries, dtype=dtype)
  pay_t = np.asarray(fixed_leg_payment_times, dtype=dtype)
  dcf = np.asarray(fixed_leg_daycount_fractions, dtype=dtype)
  coupon = np.asarray(fixed_leg_coupon, dtype=dtype)
  batch_shape = expiries.shape
  m = pay_t.shape[-1]
  exp_b = np.repeat(expiries[..., None], m, axis=-1)          # batch + [m]
  model = HullWhiteModel1F(mean_reversion, volatility, reference_rate_fn, dtype)

  from oracle import grid as grid_lib
  sim_times = np.unique(exp_b.reshape(-1))
  longest = sim_times.max()
  sim_times = np.sort(np.concatenate(
      [sim_times, grid_lib.tf_range(time_step, longest, time_step, dtype)]),
                      kind='stable')
  tau = pay_t - exp_b
  curve_times = np.unique(tau.reshape(-1))
  p_t_tau, r_t = model.sample_discount_curve_paths(
      sim_times, curve_times, num_samples, random_type, seed, skip, path_range=path_range)
  num_samples = r_t.shape[0]           # path_range (oracle extension): a slice of the paths
  # path discount factors: dt_0 = 0 (the first interval is not discounted)
  dt = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]]).astype(dtype)
  cumul = np.cumsum(r_t[:, :, 0] * dt[None, :], axis=1)       # [N, k]
  df = np.exp(-cumul)
  sim_idx = np.searchsorted(sim_times, exp_b.reshape(-1), side='left')
  curve_idx = np.searchsorted(curve_times, tau.reshape(-1), side='left')
  payoff_df = df[:, sim_idx].reshape((num_samples,) + batch_shape + (m,))
  payoff_bond = p_t_tau[:, curve_idx, sim_idx, 0].reshape(
      (num_samples,) + batch_shape + (m,))
  fixed_leg_pv = (coupon * dcf * payoff_bond).sum(axis=-1)
  float_leg_pv = 1.0 - payoff_bond[..., -1]
  payoff_swap = payoff_df[..., -1] * (float_leg_pv - fixed_leg_pv)
  payoff_swap = np.where(is_payer_swaption, payoff_swap, -payoff_swap)
  payoff = np.maximum(payoff_swap, 0.0)
  price = np.asarray(notional, dtype=dtype) * payoff.mean(axis=0)
  if return_payoffs:
    return price, payoff
  return price


def bond_option_price_mc(*, strikes, expiries, maturities, discount_rate_fn,
                         mean_reversion, volatility, is_call_options=True,
                         num_samples=1, random_type=None, seed=None, skip=0,
                         time_step=None, dtype=np.float64):
  """`bond_option_price(use_analytic_pricing=False)`
  (`hull_white/zero_coupon_bond_option.py:186-210`) with
  `options_price_from_samples` (`hjm/zero_coupon_bond_option_util.py:85-153`).
  All of `strikes`, `expiries`, `maturities` broadcast to `strikes.shape`."""
  dtype = np.dtype(dtype)
  strikes = np.asarray(strikes, dtype=dtype)
  shape = strikes.shape
  expiries = np.broadcast_to(np.asarray(expiries, dtype=dtype), shape)
  maturities = np.broadcast_to(np.asarray(maturities, dtype=dtype), shape)
  is_call = np.broadcast_to(np.asarray(is_call_options, dtype=bool), shape)
  model = HullWhiteModel1F(mean_reversion, volatility, discount_rate_fn, dtype)

  from oracle import grid as grid_lib
  sim_times = np.unique(expiries.reshape(-1))
  longest = sim_times.max()
  sim_times = np.unique(np.concatenate(
      [sim_times, grid_lib.tf_range(time_step, longest, time_step, dtype)]))
  tau = maturities - expiries
  curve_times = np.unique(tau.reshape(-1))
  p_t_tau, r_t = model.sample_discount_curve_paths(
      sim_times, curve_times, num_samples, random_type, seed, skip)
  dt = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]]).astype(dtype)
  df = np.cumprod(np.exp(-r_t[:, :, 0] * dt[None, :]), axis=1)   # [N, k]
  sim_idx = np.searchsorted(sim_times, expiries.reshape(-1), side='left')
  curve_idx = np.searchsorted(curve_times, tau.reshape(-1), side='left')
  payoff_df = df[:, sim_idx].reshape((num_samples,) + shape)
  bond = p_t_tau[:, curve_idx, sim_idx, 0].reshape((num_samples,) + shape)
  payoff = np.where(is_call, np.maximum(bond - strikes, 0.0),
                    np.maximum(strikes - bond, 0.0))
  return (payoff_df * payoff).mean(axis=0)


def cap_floor_price_mc(*, strikes, expiries, maturities, daycount_fractions,
                       reference_rate_fn, mean_reversion, volatility,
                       notional=1.0, is_cap=True, num_samples=1,
                       random_type=None, seed=None, skip=0, time_step=None,
                       dtype=np.float64):
  """`cap_floor_price(use_analytic_pricing=False)` (`hull_white/cap_floor.py:196-235`)."""
  dtype = np.dtype(dtype)
  strikes = np.asarray(strikes, dtype=dtype)
  expiries = np.asarray(expiries, dtype=dtype)
  dcf = np.asarray(daycount_fractions, dtype=dtype)
  is_cap = np.asarray(is_cap, dtype=bool)
  caplets = bond_option_price_mc(
      strikes=1.0 / (1.0 + dcf * strikes), expiries=expiries,
      maturities=maturities, discount_rate_fn=reference_rate_fn,
      mean_reversion=mean_reversion, volatility=volatility,
      is_call_options=~is_cap, num_samples=num_samples, random_type=random_type,
      seed=seed, skip=skip, time_step=time_step, dtype=dtype)
  caplets = np.where(np.broadcast_to(expiries, caplets.shape) < 0.0, 0.0, caplets)
  return np.sum(np.asarray(notional, dtype) * (1.0 + dcf * strikes) * caplets, axis=-1)


class VectorHullWhiteModel:
  """Correlated Hull-White factors, exact discretisation
  (`vector_hull_white.py:641-781` for dim > 1): constant mean reversions,
  constant or piecewise-constant volatilities (one `PiecewiseConstantFunc` or
  scalar per factor), constant or piecewise-constant correlation matrix.

  `initial_discount_rate_fn(t)` returns `t.shape` or `t.shape + [dim]` (numpy,
  analytic in t)."""

  def __init__(self, dim, mean_reversion, volatility, initial_discount_rate_fn,
               corr_matrix=None, dtype=np.float64):
    self.dim = int(dim)
    self.dtype = np.dtype(dtype)
    mr = np.broadcast_to(np.asarray(mean_reversion, dtype=self.dtype), (self.dim,))
    vols = volatility if isinstance(volatility, (list, tuple)) else [
        v for v in np.broadcast_to(np.asarray(volatility, dtype=self.dtype), (self.dim,))]

    def rate_i(i):
      def fn(t):
        r = np.asarray(initial_discount_rate_fn(t))
        return r[..., i] if r.ndim == np.ndim(t) + 1 else r
      return fn
    self.factors = [HullWhiteModel1F(mr[i], vols[i], rate_i(i), self.dtype)
                    for i in range(self.dim)]
    self.corr = corr_matrix

  def _corr_root(self, t):
    """Cholesky factors at times `t` ([n, dim, dim]); identity when no corr."""
    n = t.shape[0]
    if self.corr is None:
      return None
    if callable(self.corr):
      c = np.asarray(self.corr(t), dtype=self.dtype)
    else:
      c = np.broadcast_to(np.asarray(self.corr, dtype=self.dtype), (n, self.dim, self.dim))
    return np.linalg.cholesky(c)

  def sample_paths(self, times, num_samples, random_type=None, seed=None, skip=0,
                   times_grid=None, normal_draws=None):
    """Short rates [num_samples, k, dim]."""
    times = np.asarray(times, dtype=self.dtype)
    k = times.shape[0]
    if times_grid is None:
      jumps = [f.jumps for f in self.factors]
      all_times = np.sort(np.concatenate([np.zeros(1, self.dtype), times] + jumps),
                          kind='stable').astype(self.dtype)
      idx = np.searchsorted(all_times, times, side='left')
    else:
      all_times = np.asarray(times_grid, dtype=self.dtype)
      idx = np.minimum(np.searchsorted(all_times, times, side='left'), all_times.shape[0] - 1)
      d1 = all_times[idx] - times
      d2 = all_times[np.maximum(idx - 1, 0)] - times
      idx = np.where(np.abs(d2) > np.abs(d1), idx, np.maximum(idx - 1, 0))
    keep_mask = np.zeros(all_times.shape[0], dtype=bool)
    keep_mask[idx] = True
    dt = all_times[1:] - all_times[:-1]
    steps = dt.shape[0]
    if normal_draws is None:
      normal_draws = draws_lib.generate_mc_normal_draws(
          self.dim, steps, num_samples,
          draws_lib.RandomType.PSEUDO if random_type is None else random_type,
          seed=seed, dtype=self.dtype, skip=skip)               # [steps, N, dim]
    else:
      normal_draws = np.transpose(np.asarray(normal_draws, self.dtype), [1, 0, 2])
    exp_x = np.stack([f.conditional_mean_x(all_times) for f in self.factors], -1)   # [S, dim]
    var_x = np.stack([f.conditional_variance_x(all_times) for f in self.factors], -1)
    kk = np.asarray([f.k for f in self.factors], dtype=self.dtype)
    root = self._corr_root(all_times + dt.min() / 2) if steps else None

    def f0(t):
      return np.asarray([f.fwd(t) for f in self.factors], dtype=self.dtype)
    x = np.zeros((num_samples, self.dim), dtype=self.dtype)
    record = k != 1
    slots = [None] * k
    out = x + f0(all_times[0])
    if record:
      slots[0] = out
    written = int(keep_mask[0])
    i = 0
    while i < steps and written < k:
      normals = normal_draws[i]
      if root is not None:
        normals = np.einsum('ij,nj->ni', root[i], normals)
      vol = np.sqrt(np.maximum(var_x[i], 0))
      vol = np.where(vol > 0, vol, 0)
      x = np.exp(-kk * dt[i]) * x + exp_x[i] + vol * normals
      out = x + f0(all_times[i + 1])
      if record:
        slots[written] = out
      written += int(keep_mask[i + 1])
      i += 1
    if not record:
      return out[:, None, :]
    slots = [np.zeros_like(x) if s is None else s for s in slots]
    return np.transpose(np.stack(slots, 0), [1, 0, 2])


def vector_sample_discount_curve_paths(model, times, curve_times, num_samples, random_type=None,
                                       seed=None, skip=0):
  """`VectorHullWhiteModel.sample_discount_curve_paths` (`vector_hull_white.py:451-592`):
  (P(t, t + tau) [N, m, k, dim], short rates [N, k, dim]); factor d uses its own
  curve, mean reversion and y_d(t) in `_bond_reconstitution` (783-814)."""
  times = np.asarray(times, dtype=model.dtype)
  curve_times = np.asarray(curve_times, dtype=model.dtype)
  rates = model.sample_paths(times, num_samples, random_type, seed, skip)      # [N, k, dim]
  t = times[None, None, :]
  tau = curve_times[None, :, None]
  out = []
  for d, f in enumerate(model.factors):
    y_t = f.compute_yt(times)
    out.append(f.bond_reconstitution(t, t + tau, rates[:, None, :, d], y_t[None, None, :]))
  return np.stack(out, axis=-1), rates


def _unique_in_order(a):
  """`tf.unique`: distinct values in order of first appearance + inverse index."""
  _, first, inv = np.unique(a, return_index=True, return_inverse=True)
  order = np.argsort(first, kind='stable')
  rank = np.empty_like(order)
  rank[order] = np.arange(order.shape[0])
  return a[np.sort(first)], rank[inv]


def bermudan_swaption_price_mc(*, exercise_times, fixed_leg_payment_times,
                               fixed_leg_daycount_fractions, fixed_leg_coupon,
                               reference_rate_fn, mean_reversion, volatility,
                               notional=1.0, num_samples=100, random_type=None,
                               seed=None, skip=0, time_step=None, dtype=np.float64):
  """`bermudan_swaption_price(use_finite_difference=False)`
  (`hull_white/swaption.py:608-724`): LSM (quadratic basis on the short rate,
  one discount curve per path) on exact Hull-White paths.  `exercise_times`:
  batch + [E]; leg arrays: batch + [E, m]."""
  from oracle import grid as grid_lib
  from oracle import lsm as lsm_lib
  dtype = np.dtype(dtype)
  ex = np.asarray(exercise_times, dtype=dtype)
  pay_t = np.asarray(fixed_leg_payment_times, dtype=dtype)
  dcf = np.broadcast_to(np.asarray(fixed_leg_daycount_fractions, dtype=dtype), pay_t.shape)
  coupon = np.broadcast_to(np.asarray(fixed_leg_coupon, dtype=dtype), pay_t.shape)
  batch_shape = ex.shape[:-1]
  n_ex, m = ex.shape[-1], pay_t.shape[-1]
  nb = int(np.prod(batch_shape)) if batch_shape else 1
  ex_flat = ex.reshape(nb, n_ex)
  pay_flat = pay_t.reshape(nb, n_ex, m)
  model = HullWhiteModel1F(mean_reversion, volatility, reference_rate_fn, dtype)
  uniq, ex_index = _unique_in_order(ex_flat.reshape(-1))
  ex_index = ex_index.reshape(nb, n_ex)
  longest = uniq[-1]
  sim_times = np.unique(np.concatenate(
      [uniq, grid_lib.tf_range(time_step, longest, time_step, dtype)]))
  ex_b = np.repeat(ex_flat[..., None], m, axis=-1)
  tau = pay_flat - ex_b
  curve_times = np.unique(tau.reshape(-1))
  p_t_tau, r_t = model.sample_discount_curve_paths(
      sim_times, curve_times, num_samples, random_type, seed, skip)
  dt = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]]).astype(dtype)
  df = np.cumprod(np.exp(-r_t[:, :, 0] * dt[None, :]), axis=1)            # [N, k]
  sim_idx_e = np.searchsorted(sim_times, ex_b.reshape(-1), side='left')
  curve_idx = np.searchsorted(curve_times, tau.reshape(-1), side='left')
  bond = p_t_tau[:, curve_idx, sim_idx_e, 0].reshape((num_samples, nb, n_ex, m))
  fixed_pv = (coupon.reshape(nb, n_ex, m) * dcf.reshape(nb, n_ex, m) * bond).sum(axis=-1)
  payoff_swap = (1.0 - bond[..., -1]) - fixed_pv                         # [N, nb, E], payer
  sim_idx_u = np.searchsorted(sim_times, uniq, side='left')
  short_rate = r_t[:, sim_idx_u, :]                                      # [N, U, 1]
  u_count = uniq.shape[0]
  is_ex = np.zeros((nb, u_count), dtype=bool)
  tab = np.zeros((u_count, num_samples, nb), dtype=dtype)
  for b in range(nb):
    for e in range(n_ex):
      is_ex[b, ex_index[b, e]] = True
      tab[ex_index[b, e], :, b] = payoff_swap[:, b, e]

  def payoff_fn(rt, time_index):
    del rt
    return np.where(is_ex[:, time_index][None, :], np.maximum(tab[time_index], 0.0), 0.0)
  value = lsm_lib.least_square_mc(
      short_rate, np.arange(u_count), payoff_fn, lsm_lib.make_polynomial_basis(2),
      discount_factors=df[:, None, sim_idx_u], dtype=dtype)
  return (np.broadcast_to(np.asarray(notional, dtype), batch_shape).reshape(-1) *
          value).reshape(batch_shape)
