"""Monte-Carlo zero-coupon bond option pricing under one-factor Hull-White.

Drop-in for the simulation branch of
`tf_quant_finance.models.hull_white.bond_option_price`
(`models/hull_white/zero_coupon_bond_option.py:41-210`) together with
`options_price_from_samples` (`models/hjm/zero_coupon_bond_option_util.py:29-153`).

The reference materialises `[N, m, k, 1]` bond curves and gathers one entry per
option.  Here each option is one payoff slot of the fused HW1F price kernel:
  call  DF(t_e) max(P(t_e, T) - K, 0) = K DF max(P / K - 1, 0)
  put   DF(t_e) max(K - P(t_e, T), 0) = K DF max(1 - P / K, 0)
which is the kernel's one-payment `TQF_PAYOFF_HW_SWAPTION` form
`scale * max(+-DF (1 - coef P), 0)` with `coef = 1 / K`, `scale = K`.
"""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200.models import utils
from tff_b200.models.hull_white import _exact
from tff_b200.models.hull_white import one_factor
from tff_b200.models.hull_white import swaption as _swaption


def _bond_option_desc(model, step, expiry, maturity, strike, is_call):
  dt_ = model._dtype
  k = model._tables.k
  t_e = np.asarray(expiry, dtype=dt_)
  t_m = np.asarray(maturity, dtype=dt_)
  y = model._tables.y_t(np.asarray([expiry], dtype=dt_))[0]
  rate = lambda t: _exact.discount_rate(model._initial_discount_rate_fn, t, dt_)
  ln_p0_ratio = -(rate(t_m) * t_m) + rate(t_e) * t_e
  g = (1. - np.exp(-k * (t_m - t_e))) / k
  d = _lib.PayoffDesc()
  d.kind = _lib.PAYOFF_HW_SWAPTION
  d.expiry_step = int(step)
  d.num_payments = 1
  d.is_payer = int(not bool(is_call))
  d.scale = float(strike)
  d.pay_g[0] = float(g)
  d.pay_k[0] = float(ln_p0_ratio - 0.5 * y * g**2)
  d.pay_coef[0] = 1.0 / float(strike)
  return d


def bond_option_price(*,
                      strikes,
                      expiries,
                      maturities,
                      discount_rate_fn,
                      mean_reversion,
                      volatility,
                      is_call_options=True,
                      use_analytic_pricing=True,
                      num_samples=1,
                      random_type=None,
                      seed=None,
                      skip=0,
                      time_step=None,
                      dtype=None,
                      name=None,
                      return_stats=False):
  """Zero-coupon bond option prices of shape `strikes.shape` (numpy array).

  Same arguments as the reference.  `use_analytic_pricing=True` (the default)
  evaluates the closed form on the host; `False` runs the fused kernel.  Options whose expiry is negative are worth 0
  and are left out of the simulation grid.
  """
  del name
  dt_ = _tensor.infer_dtype(strikes, dtype, default=np.float32)
  strikes = _tensor.to_numpy(strikes, dt_)
  expiries = _tensor.to_numpy(expiries, dt_)
  maturities = _tensor.to_numpy(maturities, dt_)
  is_call = np.asarray(_tensor.to_numpy(is_call_options), dtype=bool)
  model = one_factor.HullWhiteModel1F(mean_reversion, volatility,
                                      discount_rate_fn, dtype=dt_)
  if use_analytic_pricing:
    # Black formula on the forward bond price: a closed form evaluated on the host
    # (zero_coupon_bond_option.py:210-304)
    if model._tables is None:
      raise ValueError('The paramerization of `mean_reversion` and/or `volatility` does not '
                       'support analytic computation of bond option variance.')
    from tff_b200.models.hull_white import _analytic  # pylint: disable=g-import-not-at-top
    return _analytic.bond_option_price(model, strikes, expiries, maturities, is_call).astype(dt_)
  if time_step is None:
    raise ValueError('`time_step` must be provided for simulation '
                     'based bond option valuation.')
  if model._tables is None:
    raise NotImplementedError(
        'bond_option_price needs constant mean reversion and constant or '
        'piecewise-constant volatility (exact discretisation).')
  shape = strikes.shape
  k_flat = strikes.reshape(-1)
  e_flat = np.broadcast_to(expiries, shape).reshape(-1)
  m_flat = np.broadcast_to(maturities, shape).reshape(-1)
  c_flat = np.broadcast_to(is_call, shape).reshape(-1)
  live = np.nonzero(e_flat >= 0)[0]
  price = np.zeros(k_flat.shape[0], dtype=dt_)
  stderr = np.zeros(k_flat.shape[0], dtype=np.float64)
  if live.size:
    # sim_times: unique expiries plus the uniform grid, de-duplicated
    # (hjm/zero_coupon_bond_option_util.py:87-94)
    sim_times = np.unique(e_flat[live])
    longest = sim_times.max()
    sim_times = np.unique(np.concatenate(
        [sim_times, utils._tf_range(time_step, longest, time_step, dt_)])).astype(dt_)

    def make_desc(b, step):
      i = live[b]
      return _bond_option_desc(model, step, e_flat[i], m_flat[i], k_flat[i], c_flat[i])
    sums, n = _swaption._price_on_grid(model, sim_times, e_flat[live], make_desc,
                                       num_samples, random_type, seed, skip)
    price[live] = (sums[:, 0] / n).astype(dt_)
    var = np.maximum(sums[:, 1] / n - (sums[:, 0] / n)**2, 0.0)
    stderr[live] = np.sqrt(var / n)
  price = price.reshape(shape)
  if return_stats:
    return price, stderr.reshape(shape)
  return price
