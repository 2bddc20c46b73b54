"""Oracle (test infrastructure): Gaussian / quasi-Gaussian HJM models and the
Monte-Carlo swaption price on them.

Restates, op for op in numpy on top of `oracle.euler.sample`:
  * `models/hjm/quasi_gaussian_hjm.py`: the closures 232-289 (state `[x, vec(y)]`
    of dimension F + F^2, volatility padded with zero rows/columns -- so every
    Euler step CONSUMES F + F^2 normal draws of which only the first F act),
    `_sample_paths` 451-497 (the grid of `prepare_grid` is handed to
    `euler_sampling.sample`, which grids it again; rate = sum x + f(0, t); discount
    factor exp(-cumsum(r_j (t_j - t_{j-1}))), i.e. the RIGHT-point rule),
    `sample_discount_curve_paths` 365-449, `_bond_reconstitution` 499-525 and
    `_get_valid_sqrt_matrix` 528-545;
  * `models/hjm/gaussian_hjm.py`: closures 203-228 (state x only, y(t) in closed
    form), `state_y` 316-373, `discount_bond_price` 375-411, `_sample_paths`
    413-461 (discount factor cumprod(exp(-r_i (t_{i+1} - t_i))): the LEFT-point rule);
  * `models/hjm/swaption_pricing.py:341-392` (`_european_swaption_mc`) with
    `models/hjm/swaption_util.py:28-170`.
Pinned by the reference's own values: `gaussian_hjm_test.py:224-283` (bond prices,
1e-8), `swaption_pricing_test.py:46-127, 321-356` (0.71632434 / 0.81348254 /
0.802226 to 1e-2).  `initial_discount_rate_fn` must be analytic in t (complex-step
forward rate, as `oracle/hull_white.py`).  A `volatility` callable takes numpy
`(t, r_t)` with `r_t` of shape `[num_samples, 1]`.
"""
import numpy as np

from oracle import euler as euler_lib
from oracle import grid as grid_lib
from oracle import models as models_lib


def _valid_sqrt_matrix(rho):
  """`_get_valid_sqrt_matrix` (quasi_gaussian_hjm.py:528-545)."""
  e, v = np.linalg.eigh(rho)
  if np.any(e < 1e-5):
    return v @ np.sqrt(np.diag(np.maximum(e, 1e-5)))
  return np.linalg.cholesky(rho)


class QuasiGaussianHJM:
  """`QuasiGaussianHJM` without model batching."""

  def __init__(self, dim, mean_reversion, volatility, initial_discount_rate_fn,
               corr_matrix=None, dtype=np.float64):
    self.dtype = np.dtype(dtype)
    self.factors = int(dim)
    self.dim = self.factors + self.factors**2
    self.k = np.asarray(mean_reversion, dtype=self.dtype).reshape(self.factors)
    self.rate_fn = initial_discount_rate_fn
    self.fwd, _ = models_lib.complex_step_forward_rate(initial_discount_rate_fn)
    if callable(volatility):
      self.vol_fn = volatility
    else:
      v = np.asarray(volatility, dtype=self.dtype).reshape(self.factors)
      self.vol_fn = lambda t, r: v
    rho = np.eye(self.factors, dtype=self.dtype) if corr_matrix is None else np.asarray(
        corr_matrix, dtype=self.dtype)
    self.rho = rho
    self.sqrt_rho = _valid_sqrt_matrix(rho).astype(self.dtype)

  # closures (232-289)
  def _vol(self, t, x):
    r_t = self.dtype.type(self.fwd(t)) + x.sum(-1, keepdims=True)
    return np.broadcast_to(np.asarray(self.vol_fn(t, r_t), dtype=self.dtype),
                           x.shape[:-1] + (self.factors,))

  def volatility_fn(self, t, state):
    f = self.factors
    x = state[..., :f]
    vol = self._vol(t, x)[..., None]                          # [N, F, 1]
    out = np.zeros(state.shape[:-1] + (self.dim, self.dim), dtype=self.dtype)
    out[..., :f, :f] = self.sqrt_rho * vol
    return out

  def drift_fn(self, t, state):
    f = self.factors
    x = state[..., :f]
    y = state[..., f:].reshape(state.shape[:-1] + (f, f))
    vol = self._vol(t, x)[..., None]
    vol_sq = vol @ np.swapaxes(vol, -1, -2)
    mr2 = self.k[:, None] + self.k[None, :]
    drift_x = y.sum(-1) - self.k * x
    drift_y = (self.rho * vol_sq - mr2 * y).reshape(state.shape[:-1] + (f * f,))
    return np.concatenate([drift_x, drift_y], -1)

  def _sample_paths(self, times, time_step, num_time_steps, num_samples, random_type, skip, seed):
    dt_ = self.dtype
    times = np.asarray(times, dtype=dt_)
    ts_internal = time_step
    if num_time_steps is not None:
      ts_internal = dt_.type(times[-1] / dt_.type(num_time_steps))
    grid, _, idx = grid_lib.prepare_grid(times=times, time_step=ts_internal, dtype=dt_,
                                         num_time_steps=num_time_steps)
    dt = grid[1:] - grid[:-1]
    xy = euler_lib.sample(self.dim, self.drift_fn, self.volatility_fn, grid,
                          num_samples=num_samples, initial_state=np.zeros(self.dim, dt_),
                          random_type=random_type, seed=seed, time_step=time_step,
                          num_time_steps=num_time_steps, skip=skip, dtype=dt_)
    x = xy[..., :self.factors]
    y = xy[..., self.factors:]
    f0 = np.asarray(self.fwd(grid), dtype=dt_)
    rate = x.sum(-1) + f0[None, :]
    dts = np.concatenate([np.zeros(1, dt_), dt])
    df = np.exp(-np.cumsum(rate * dts, axis=-1))
    return rate[:, idx], df[:, idx], x[:, idx], y[:, idx]

  def sample_paths(self, times, num_samples, time_step=None, num_time_steps=None,
                   random_type=None, seed=None, skip=0):
    return self._sample_paths(times, time_step, num_time_steps, num_samples, random_type, skip,
                              seed)

  def _bond_reconstitution(self, times, maturities, x_t, y_t):
    """Eq. 10.18 (499-525): times [1,1,k], maturities [1,m,k], x_t [N,1,k,F],
    y_t [N,1,k,F,F] -> [N,m,k]."""
    te, me = times[..., None], maturities[..., None]
    p0t = np.exp(-self.rate_fn(times) * times)
    p0 = np.exp(-self.rate_fn(maturities) * maturities) / p0t
    g = (1. - np.exp(-self.k * (me - te))) / self.k                     # [1,m,k,F]
    term1 = (x_t * g).sum(-1)
    term2 = (g * np.einsum('...ij,...j->...i', y_t, g)).sum(-1)
    return p0 * np.exp(-term1 - 0.5 * term2)

  def sample_discount_curve_paths(self, times, curve_times, num_samples, time_step=None,
                                  num_time_steps=None, random_type=None, seed=None, skip=0):
    dt_ = self.dtype
    times = np.asarray(times, dtype=dt_)
    curve_times = np.asarray(curve_times, dtype=dt_)
    rate, df, x_t, y_t = self._sample_paths(times, time_step, num_time_steps, num_samples,
                                            random_type, skip, seed)
    f = self.factors
    x_t = x_t[:, None]                                                   # [N,1,k,F]
    y_t = y_t.reshape(y_t.shape[0], 1, times.shape[0], f, f)
    t3 = times.reshape(1, 1, -1)
    c3 = curve_times.reshape(1, -1, 1)
    return self._bond_reconstitution(t3, t3 + c3, x_t, y_t), rate, df


class GaussianHJM(QuasiGaussianHJM):
  """`GaussianHJM`: deterministic volatility (constant or piecewise constant per factor)."""

  def __init__(self, dim, mean_reversion, volatility, initial_discount_rate_fn,
               corr_matrix=None, dtype=np.float64):
    dtype = np.dtype(dtype)
    if isinstance(volatility, models_lib.PiecewiseConstantFunc):
      self.vol_pw = volatility                       # jumps [F, J] or [J]; values [F, J+1]
    else:
      self.vol_pw = None
      self.vol_const = np.asarray(volatility, dtype=dtype).reshape(int(dim))
    super().__init__(dim, mean_reversion, lambda t, r: self._sigma(t), initial_discount_rate_fn,
                     corr_matrix, dtype)
    self.dim = self.factors                          # state x only (gaussian_hjm.py:161-162)
    self.sqrt_rho = np.linalg.cholesky(self.rho).astype(self.dtype)     # :203

  def _jumps_values(self):
    f = self.factors
    if self.vol_pw is None:
      return np.zeros((f, 0), self.dtype), self.vol_const.reshape(f, 1)
    j = np.asarray(self.vol_pw.jump_locations(), dtype=self.dtype)
    v = np.asarray(self.vol_pw.values(), dtype=self.dtype)
    j = np.broadcast_to(j.reshape(-1, j.shape[-1]), (f, j.shape[-1]))
    v = np.broadcast_to(v.reshape(-1, v.shape[-1]), (f, v.shape[-1]))
    return j, v

  def _sigma(self, t):
    """sigma_i(t), left-continuous (PiecewiseConstantFunc): [..., F] for t [...]."""
    jumps, values = self._jumps_values()
    t = np.asarray(t, dtype=self.dtype)
    out = np.empty(t.shape + (self.factors,), dtype=self.dtype)
    for i in range(self.factors):
      out[..., i] = values[i][np.searchsorted(jumps[i], t, side='left')]
    return out

  def state_y(self, t):
    """y_ij(t) = e^{-(k_i+k_j) t} int_0^t rho_ij sigma_i sigma_j e^{(k_i+k_j) u} du
    (316-373) -> [F, F, len(t)]."""
    t = np.asarray(t, dtype=self.dtype).reshape(-1)
    jumps, values = self._jumps_values()
    f = self.factors
    mr2 = self.k[:, None] + self.k[None, :]
    out = np.zeros((f, f, t.shape[0]), dtype=self.dtype)
    for i in range(f):
      for j in range(f):
        knots = np.unique(np.concatenate([jumps[i], jumps[j]]))
        for n, tt in enumerate(t):
          edges = np.concatenate([[0.0], knots[knots < tt], [tt]])
          acc = 0.0
          for a, b in zip(edges[:-1], edges[1:]):
            mid = 0.5 * (a + b)
            si = values[i][np.searchsorted(jumps[i], mid, side='left')]
            sj = values[j][np.searchsorted(jumps[j], mid, side='left')]
            acc += self.rho[i, j] * si * sj / mr2[i, j] * (np.exp(mr2[i, j] * b) - np.exp(mr2[i, j] * a))
          out[i, j, n] = np.exp(-mr2[i, j] * tt) * acc
    return out

  def volatility_fn(self, t, state):
    vol = self._sigma(t).reshape(self.factors, 1)
    return np.broadcast_to(self.sqrt_rho * vol, state.shape[:-1] + (self.factors, self.factors))

  def drift_fn(self, t, state):
    y = self.state_y(np.asarray([t]))[..., 0]
    return y.sum(-1) - self.k * state

  def discount_bond_price(self, state, times, maturities):
    """375-411: state [n, F], times [n], maturities [n] -> [n]."""
    x_t = np.asarray(state, dtype=self.dtype)
    times = np.asarray(times, dtype=self.dtype)
    maturities = np.asarray(maturities, dtype=self.dtype)
    y_t = np.transpose(self.state_y(times)).reshape(times.shape + (self.factors, self.factors))
    return self._bond_reconstitution(times, maturities, x_t, y_t)

  def _sample_paths(self, times, time_step, num_time_steps, num_samples, random_type, skip, seed):
    dt_ = self.dtype
    times = np.asarray(times, dtype=dt_)
    ts_internal = time_step
    if num_time_steps is not None:
      ts_internal = dt_.type(times[-1] / dt_.type(num_time_steps))
    grid, _, idx = grid_lib.prepare_grid(times=times, time_step=ts_internal, dtype=dt_,
                                         num_time_steps=num_time_steps)
    dt = grid[1:] - grid[:-1]
    x = euler_lib.sample(self.dim, self.drift_fn, self.volatility_fn, grid,
                         num_time_steps=num_time_steps, num_samples=num_samples,
                         initial_state=np.zeros(self.dim, dt_), random_type=random_type,
                         seed=seed, time_step=time_step, skip=skip, dtype=dt_)
    y = self.state_y(grid).reshape(self.factors**2, -1).T          # [times, F^2]
    y = np.broadcast_to(y[None], (num_samples,) + y.shape)
    f0 = np.asarray(self.fwd(grid), dtype=dt_)
    rate = x.sum(-1) + f0[None, :]
    df = np.exp(-rate[:, :-1] * dt)
    df = np.cumprod(np.concatenate([np.ones((num_samples, 1), dt_), df], axis=1), axis=1)
    return rate[:, idx], df[:, idx], x[:, idx], y[:, idx]


def swaption_price_mc(*, expiries, fixed_leg_payment_times, fixed_leg_daycount_fractions,
                      fixed_leg_coupon, reference_rate_fn, num_hjm_factors, mean_reversion,
                      volatility, time_step=None, num_time_steps=None, corr_matrix=None,
                      notional=1.0, is_payer_swaption=True, num_samples=1, random_type=None,
                      seed=None, skip=0, dtype=np.float64, return_payoffs=False):
  """`hjm.swaption_price` (MONTE_CARLO): `swaption_pricing.py:255-392` with
  `swaption_util.py:94-170` (sim_times = sorted expiries, curve_times = unique
  payment_time - expiry).  Swaption batch shape = expiries.shape."""
  dtype = np.dtype(dtype)
  expiries = np.asarray(expiries, dtype=dtype)
  pay = np.asarray(fixed_leg_payment_times, dtype=dtype)
  dcf = np.broadcast_to(np.asarray(fixed_leg_daycount_fractions, dtype=dtype), pay.shape)
  cpn = np.broadcast_to(np.asarray(fixed_leg_coupon, dtype=dtype), pay.shape)
  model = QuasiGaussianHJM(num_hjm_factors, mean_reversion, volatility, reference_rate_fn,
                           corr_matrix, dtype)
  exp_rep = np.repeat(expiries[..., None], pay.shape[-1], axis=-1)           # batch + [m]
  pay, dcf, cpn = (np.broadcast_to(a, exp_rep.shape) for a in (pay, dcf, cpn))
  sim_times = np.sort(exp_rep.reshape(-1))
  tau = pay - exp_rep
  curve_times = np.unique(tau.reshape(-1))
  p_t_tau, _, df = model.sample_discount_curve_paths(
      sim_times, curve_times, num_samples, time_step=time_step, num_time_steps=num_time_steps,
      random_type=random_type, seed=seed, skip=skip)
  sim_idx = np.searchsorted(sim_times, exp_rep.reshape(-1))
  cur_idx = np.searchsorted(curve_times, tau.reshape(-1))
  bond = p_t_tau[:, cur_idx, sim_idx].reshape((num_samples,) + pay.shape)    # [N, batch, m]
  dfs = df[:, sim_idx].reshape((num_samples,) + pay.shape)
  fixed = (cpn * dcf * bond).sum(-1)
  swap = dfs[..., -1] * ((1.0 - bond[..., -1]) - fixed)
  swap = np.where(np.asarray(is_payer_swaption, dtype=bool), swap, -swap)
  payoff = np.maximum(swap, 0.0)
  price = np.asarray(notional, dtype=dtype) * payoff.mean(0)
  return (price, payoff) if return_payoffs else price


def bond_option_price_mc(*, strikes, expiries, maturities, discount_rate_fn, dim, mean_reversion,
                         volatility, corr_matrix=None, is_call_options=True, num_samples=1,
                         random_type=None, seed=None, skip=0, time_step=None, dtype=np.float64):
  """`hjm.bond_option_price` (`zero_coupon_bond_option.py:30-195`) with
  `options_price_from_samples` (`zero_coupon_bond_option_util.py:29-153`):
  sim_times = unique(expiries U range(time_step, longest, time_step)); the discount
  factor is cumprod(exp(-r(sim_i) (sim_i - sim_{i-1}))) over the SIM times with
  dt_0 = 0 (the model's own discount factors are dropped, line 170)."""
  if time_step is None:
    raise ValueError('`time_step` must be provided for simulation based bond option valuation.')
  dtype = np.dtype(dtype)
  strikes = np.asarray(strikes, dtype=dtype)
  expiries = np.broadcast_to(np.asarray(expiries, dtype=dtype), strikes.shape)
  maturities = np.broadcast_to(np.asarray(maturities, dtype=dtype), strikes.shape)
  is_call = np.broadcast_to(np.asarray(is_call_options, dtype=bool), strikes.shape)
  model = QuasiGaussianHJM(dim, mean_reversion, volatility, discount_rate_fn, corr_matrix, dtype)
  sim_times = np.unique(expiries.reshape(-1))
  longest = sim_times.max()
  sim_times = np.unique(np.concatenate(
      [sim_times, grid_lib.tf_range(dtype.type(time_step), longest, dtype.type(time_step), dtype)]))
  tau = maturities - expiries
  curve_times = np.unique(tau.reshape(-1))
  p_t_tau, r_t, _ = model.sample_discount_curve_paths(
      sim_times, curve_times, num_samples, time_step=time_step, random_type=random_type,
      seed=seed, skip=skip)
  dt = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]]).astype(dtype)
  df = np.cumprod(np.exp(-r_t * dt[None, :]), axis=1)
  sim_idx = np.searchsorted(sim_times, expiries.reshape(-1))
  cur_idx = np.searchsorted(curve_times, tau.reshape(-1))
  bond = p_t_tau[:, cur_idx, sim_idx].reshape((num_samples,) + strikes.shape)
  dfs = df[:, sim_idx].reshape((num_samples,) + strikes.shape)
  payoff = np.where(is_call, np.maximum(bond - strikes, 0.0), np.maximum(strikes - bond, 0.0))
  return (dfs * payoff).mean(0)


def cap_floor_price_mc(*, strikes, expiries, maturities, daycount_fractions, reference_rate_fn, dim,
                       mean_reversion, volatility, corr_matrix=None, notional=1.0, is_cap=True,
                       num_samples=1, random_type=None, seed=None, skip=0, time_step=None,
                       dtype=np.float64):
  """`hjm.cap_floor_price` (`cap_floor.py:27-229`)."""
  dtype = np.dtype(dtype)
  strikes = np.asarray(strikes, dtype=dtype)
  expiries = np.asarray(expiries, dtype=dtype)
  dcf = np.asarray(daycount_fractions, dtype=dtype)
  caplets = bond_option_price_mc(
      strikes=1.0 / (1.0 + dcf * strikes), expiries=expiries, maturities=maturities,
      discount_rate_fn=reference_rate_fn, dim=dim, mean_reversion=mean_reversion,
      volatility=volatility, corr_matrix=corr_matrix,
      is_call_options=~np.asarray(is_cap, dtype=bool), num_samples=num_samples,
      random_type=random_type, seed=seed, skip=skip, time_step=time_step, dtype=dtype)
  caplets = np.where(np.broadcast_to(expiries, caplets.shape) < 0.0, 0.0, caplets)
  return (np.asarray(notional, dtype=dtype) * (1.0 + dcf * strikes) * caplets).sum(-1)
