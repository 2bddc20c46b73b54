"""Monte-Carlo zero-coupon bond option price under the HJM model
(`models/hjm/zero_coupon_bond_option.py:30-195` with `options_price_from_samples`,
`models/hjm/zero_coupon_bond_option_util.py:29-153`).

The reference samples `[N, m, k]` bond curves on the grid of the expiries and the
uniform `time_step` grid, gathers one entry per option and discounts with
`DF(t_j) = prod_{i <= j} exp(-r(t_i) (t_i - t_{i-1}))` over those SIMULATION times
(`dt_0 = 0`; the model's own discount factors are dropped, line 170).  Here every
option is one payoff slot of the fused HJM path kernel,
  call  DF max(P - K, 0) = K DF max(P / K - 1, 0),  put  K DF max(1 - P / K, 0),
the one-payment form `scale max(+-DF (1 - coef P), 0)` of `TQF_PAYOFF_HW_SWAPTION`
(`coef = 1 / K`, `scale = K`, `P = exp(K_0 - G . x)`), and the short-rate integral
the kernel carries is given the reference's weights; nothing is stored.
"""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import distributed
from tff_b200.models import utils
from tff_b200.models.hjm import quasi_gaussian_hjm
from tff_b200.models.hjm.swaption_pricing import _RawPayoff


def _bond_option_desc(model, entry, y_e, expiry, maturity, strike, is_call):
  dt_ = model._dtype
  f = model._factors
  a, g = model._bond_tables(np.asarray([expiry], dtype=dt_),
                            np.asarray([[maturity]], dtype=dt_), y_e[None])   # [1, 1], [1, 1, F]
  d = _lib.PayoffDesc()
  d.kind = _lib.PAYOFF_HW_SWAPTION
  d.expiry_step = int(entry)
  d.num_payments = 1
  d.num_factors = f
  d.is_payer = int(not bool(is_call))
  d.scale = float(strike)
  d.pay_k[0] = float(np.log(a[0, 0]))
  d.pay_coef[0] = 1.0 / float(strike)
  for i in range(f):
    d.pay_g[i] = float(g[0, 0, i])
  return d


def _simulation_grid(expiries, time_step, dt_):
  """(sim_times, integral_weights_fn) of `options_price_from_samples`
  (zero_coupon_bond_option_util.py:87-113): the unique expiries merged with the uniform
  grid; the discount factor sums `r(t_j) (t_j - t_{j-1})` over THOSE times, `dt_0 = 0`."""
  sim_times = np.unique(expiries)
  longest = sim_times.max()
  sim_times = np.unique(np.concatenate(
      [sim_times, utils._tf_range(time_step, longest, time_step, dt_)])).astype(dt_)   # pylint: disable=protected-access

  def integral_weights(all_times, entries):
    # the Euler step that lands on simulation time j >= 1 carries t_j - t_{j-1}
    w = np.zeros(all_times.shape[0] - 1, dtype=dt_)
    dts = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]]).astype(dt_)
    for j, e in enumerate(entries):
      if e >= 1:
        w[e - 1] += dts[j]
    return w
  return sim_times, integral_weights


def bond_option_price(*, strikes, expiries, maturities, discount_rate_fn, dim, mean_reversion,
                      volatility, corr_matrix=None, is_call_options=True, num_samples=1,
                      random_type=None, seed=None, skip=0, time_step=None, dtype=None, name=None,
                      return_stats=False):
  """`tff.models.hjm.bond_option_price`: prices of shape `strikes.shape` (numpy).
  Same arguments as the reference."""
  del name
  if time_step is None:
    raise ValueError('`time_step` must be provided for simulation based '
                     'bond option valuation.')
  dt_ = _tensor.infer_dtype(strikes, dtype, default=np.float32)
  strikes = _tensor.to_numpy(strikes, dt_)
  shape = strikes.shape
  k_flat = strikes.reshape(-1)
  e_flat = np.broadcast_to(_tensor.to_numpy(expiries, dt_), shape).reshape(-1)
  m_flat = np.broadcast_to(_tensor.to_numpy(maturities, dt_), shape).reshape(-1)
  c_flat = np.broadcast_to(np.asarray(_tensor.to_numpy(is_call_options), dtype=bool), shape).reshape(-1)
  model = quasi_gaussian_hjm.QuasiGaussianHJM(
      dim, mean_reversion=mean_reversion, volatility=volatility,
      initial_discount_rate_fn=discount_rate_fn, corr_matrix=corr_matrix, dtype=dt_)
  ts = dt_.type(_tensor.to_numpy(time_step))
  sim_times, integral_weights = _simulation_grid(e_flat, ts, dt_)

  plan, _, entry_of, inverse, y_entries, sim_times = model._plan(   # pylint: disable=protected-access
      sim_times, ts, None, num_samples, random_type, seed, skip,
      integral_weights_fn=integral_weights)
  n = float(plan.num_samples)
  price = np.zeros(k_flat.shape[0], dtype=np.float64)
  second = np.zeros(k_flat.shape[0], dtype=np.float64)
  bad = np.zeros(k_flat.shape[0], dtype=np.float64)
  try:
    sim_idx = np.searchsorted(sim_times, e_flat, side='left')
    descs, slots = [], []
    for b in range(e_flat.shape[0]):
      entry = int(entry_of[inverse[sim_idx[b]]])
      y_e = model._y_at(sim_times[sim_idx[b]:sim_idx[b] + 1], y_entries[entry][None])[0]   # pylint: disable=protected-access
      if entry == 0:
        # expiry on the start of the grid: x = 0, DF = 1 -- a deterministic payoff
        a, _ = model._bond_tables(np.asarray([e_flat[b]], dtype=dt_),   # pylint: disable=protected-access
                                  np.asarray([[m_flat[b]]], dtype=dt_), y_e[None])
        v = max(a[0, 0] - k_flat[b], 0.0) if c_flat[b] else max(k_flat[b] - a[0, 0], 0.0)
        price[b], second[b] = v, v * v
        continue
      descs.append(_RawPayoff(_bond_option_desc(model, entry, y_e, e_flat[b], m_flat[b], k_flat[b],
                                                c_flat[b])))
      slots.append(b)
    for c0 in range(0, len(descs), _lib.MAX_PAYOFFS):
      sums = distributed.price_sums_host(plan, descs[c0:c0 + _lib.MAX_PAYOFFS])
      for r, b in enumerate(slots[c0:c0 + _lib.MAX_PAYOFFS]):
        price[b], second[b], bad[b] = sums[r, 0] / n, sums[r, 1] / n, sums[r, 2]
  finally:
    plan.close()
  out = price.astype(dt_).reshape(shape)
  if not return_stats:
    return out
  var = np.maximum(second - price**2, 0.0)
  return out, np.sqrt(var / n).reshape(shape), bad.reshape(shape)
