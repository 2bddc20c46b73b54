"""Closed-form Hull-White valuations on the host (numpy):
`_analytic_valuation` of `hull_white/zero_coupon_bond_option.py:210-304` (Black
formula on the forward bond price with the model's bond-option variance) and of
`hull_white/swaption.py:726-814, 937-984` (Jamshidian decomposition: a swaption
is a portfolio of bond options struck at the bond prices of the break-even short
rate).  They are what the reference's pricers return by default
(`use_analytic_pricing=True`); no Monte-Carlo path is involved, so nothing runs
on the GPU.  Constant mean reversion, constant or piecewise-constant volatility.
"""
import numpy as np
from scipy import optimize
from scipy import special

from tff_b200.models.hull_white import _exact


def _ncdf(x):
  return 0.5 * (1.0 + special.erf(x / np.sqrt(2.0)))


def _rate(model, t):
  return _exact.discount_rate(model._initial_discount_rate_fn, t, model._dtype)


def bond_option_variance(model, expiries, maturities):
  """Black-equivalent variance of `P(T0, T)` (`_bond_option_variance` 252-304):
  y(T0) G(T - T0)^2 with y(t) = e^{-2 a t} int_0^t sigma^2 e^{2 a s} ds."""
  k = model._tables.k
  y = model._tables.y_t(expiries.reshape(-1)).reshape(expiries.shape)
  g = (1.0 - np.exp(-k * (maturities - expiries))) / k
  return y * g**2


def bond_option_price(model, strikes, expiries, maturities, is_call):
  """`_analytic_valuation` (`zero_coupon_bond_option.py:210-248`)."""
  shape = np.broadcast(strikes, expiries, maturities).shape
  strikes, expiries, maturities, is_call = (np.broadcast_to(a, shape) for a in (
      strikes, expiries, maturities, is_call))
  df_e = np.exp(-_rate(model, expiries) * expiries)
  df_m = np.exp(-_rate(model, maturities) * maturities)
  variance = bond_option_variance(model, expiries, maturities)
  fwd = df_m / df_e
  sq = np.sqrt(variance)
  with np.errstate(all='ignore'):
    d1 = np.where(sq > 0, (np.log(fwd / strikes) + 0.5 * variance) / np.where(sq > 0, sq, 1.0), 0.0)
  d2 = d1 - sq
  call = df_m * _ncdf(d1) - strikes * df_e * _ncdf(d2)
  put = strikes * df_e * _ncdf(-d2) - df_m * _ncdf(-d1)
  intrinsic = np.where(is_call, np.maximum(fwd - strikes, 0), np.maximum(strikes - fwd, 0))
  value = np.where(sq > 0.0, np.where(is_call, call, put), intrinsic)
  return np.where(maturities < expiries, 0.0, value)


def _bond_price_given_rate(model, r, expiry, maturities):
  """P(T0, T | r(T0) = r) (`_bond_reconstitution`, vector_hull_white.py:783-814)."""
  k = model._tables.k
  e = np.asarray([expiry], dtype=model._dtype)
  y = model._tables.y_t(e)[0]
  g = (1.0 - np.exp(-k * (maturities - expiry))) / k
  p0 = np.exp(-_rate(model, maturities) * maturities) / np.exp(-_rate(model, e)[0] * expiry)
  x = r - np.asarray(model._fwd(e))[0]
  return p0 * np.exp(-x * g - 0.5 * y * g**2)


def swaption_price(model, expiries, pay_times, dcf, coupon, notional, is_payer):
  """`_analytic_valuation` (`swaption.py:937-984`) for `expiries` of shape
  `batch` and leg arrays of shape `batch + [m]`."""
  batch_shape = expiries.shape
  m = pay_times.shape[-1]
  exp_f = expiries.reshape(-1)
  pay_f = np.broadcast_to(pay_times, batch_shape + (m,)).reshape(-1, m)
  coef_f = (np.broadcast_to(dcf, batch_shape + (m,)) *
            np.broadcast_to(coupon, batch_shape + (m,))).reshape(-1, m)
  payer_f = np.broadcast_to(is_payer, batch_shape).reshape(-1)
  out = np.zeros(exp_f.shape[0], dtype=np.float64)
  for b in range(exp_f.shape[0]):
    jam = np.concatenate([-coef_f[b, :-1], [-1.0 - coef_f[b, -1]]])

    def zero_fun(r, b=b, jam=jam):
      return float(np.sum(jam * _bond_price_given_rate(model, r, exp_f[b], pay_f[b])) + 1.0)
    r_star = optimize.brentq(zero_fun, -1.0, 1.0, xtol=1e-14, rtol=1e-14)
    strikes = _bond_price_given_rate(model, r_star, exp_f[b], pay_f[b])
    # payer swaption = portfolio of bond puts (is_call = not payer)
    opts = bond_option_price(model, strikes, np.full(m, exp_f[b]), pay_f[b],
                             np.full(m, not payer_f[b]))
    out[b] = np.sum(opts * coef_f[b]) + opts[-1]
  return np.broadcast_to(notional, batch_shape) * out.reshape(batch_shape)
