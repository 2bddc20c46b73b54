"""Monte-Carlo European swaption price under the HJM model
(`models/hjm/swaption_pricing.py:36-392` with `models/hjm/swaption_util.py:28-170`).

The reference simulates the quasi-Gaussian state on the grid of the sorted
expiries, materialises the bond curves `[N, m, k]`, gathers them at the payoff
times and reduces.  Here each swaption is ONE payoff descriptor of the fused path
kernel, evaluated in registers when the path reaches the Euler entry its expiry is
read at:
  payoff = notional max(+-DF(t_e) (1 - P_N - sum_j c_j tau_j P_j), 0),
  P_j = exp(K_j - sum_i G_ji x_i),  DF = exp(-I)
(`TQF_PAYOFF_HW_SWAPTION` with `num_factors = F` on `TQF_MODEL_HJM`); nothing is
stored.  The finite-difference valuation method is a PDE solver outside the
Monte-Carlo hot path and is not provided.
"""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import distributed
from tff_b200.models.hjm import quasi_gaussian_hjm
from tff_b200.models.hull_white import _exact


class _RawPayoff:
  def __init__(self, d):
    self._d = d

  def desc(self):
    return self._d


def _swaption_desc(model, entry, y_e, expiry, pay_times, coupon, dcf, is_payer, notional):
  dt_ = model._dtype
  f = model._factors
  n = int(pay_times.shape[0])
  if n * f > _lib.MAX_SWAPTION_PAYMENTS:
    raise NotImplementedError('at most {} fixed-leg payments per swaption for {} factors'.format(
        _lib.MAX_SWAPTION_PAYMENTS // f, f))
  a, g = model._bond_tables(np.asarray([expiry], dtype=dt_), pay_times.reshape(n, 1),
                            y_e[None])                       # [n, 1], [n, 1, F]
  d = _lib.PayoffDesc()
  d.kind = _lib.PAYOFF_HW_SWAPTION
  d.expiry_step = int(entry)
  d.num_payments = n
  d.num_factors = f
  d.is_payer = int(bool(is_payer))
  d.scale = float(notional)
  coef = np.array(coupon * dcf, dtype=np.float64)
  coef[-1] += 1.0                                            # float leg: 1 - P(t_e, T_N)
  for j in range(n):
    d.pay_k[j] = float(np.log(a[j, 0]))
    d.pay_coef[j] = float(coef[j])
    for i in range(f):
      d.pay_g[j * f + i] = float(g[j, 0, i])
  return d


def price(*, expiries, fixed_leg_payment_times, fixed_leg_daycount_fractions, fixed_leg_coupon,
          reference_rate_fn, num_hjm_factors, mean_reversion, volatility, times=None,
          time_step=None, num_time_steps=None, curve_times=None, corr_matrix=None, notional=None,
          is_payer_swaption=None, valuation_method=None, num_samples=1, random_type=None,
          seed=None, skip=0, time_step_finite_difference=None,
          num_time_steps_finite_difference=None, num_grid_points_finite_difference=101,
          dtype=None, name=None, return_stats=False):
  """`tff.models.hjm.swaption_price`: prices of shape `expiries.shape` (numpy).

  Same arguments as the reference.  `valuation_method` None / MONTE_CARLO only;
  `times` (custom simulation times) must contain the expiries and `curve_times` is
  not needed (the bond prices are evaluated in-kernel at the payment times)."""
  del name, curve_times, time_step_finite_difference, num_time_steps_finite_difference
  del num_grid_points_finite_difference
  if valuation_method is not None and getattr(valuation_method, 'name', str(valuation_method)) not in (
      'MONTE_CARLO', 'ValuationMethod.MONTE_CARLO'):
    raise NotImplementedError(
        'The finite-difference swaption valuation is a PDE solver outside the B200 '
        'Monte-Carlo hot path; use valuation_method=MONTE_CARLO.')
  dt_ = _tensor.infer_dtype(expiries, dtype, default=np.float32)
  expiries = _tensor.to_numpy(expiries, dt_)
  pay_t = _tensor.to_numpy(fixed_leg_payment_times, dt_)
  dcf = np.broadcast_to(_tensor.to_numpy(fixed_leg_daycount_fractions, dt_), pay_t.shape)
  coupon = np.broadcast_to(_tensor.to_numpy(fixed_leg_coupon, dt_), pay_t.shape)
  if expiries.ndim < pay_t.ndim - 1:
    raise ValueError('Swaption expiries not specified for all swaptions '
                     'in the batch. Expected rank {} but received {}.'.format(
                         pay_t.ndim - 1, expiries.ndim))
  if times is None and time_step is None and num_time_steps is None:
    raise ValueError('One of `times`, `time_step` or `num_time_steps` must be '
                     'provided for simulation based swaption valuation.')
  ntl = np.asarray(1.0 if notional is None else _tensor.to_numpy(notional, dt_), dtype=dt_)
  payer = np.asarray(True if is_payer_swaption is None else _tensor.to_numpy(is_payer_swaption),
                     dtype=bool)
  model = quasi_gaussian_hjm.QuasiGaussianHJM(
      num_hjm_factors, mean_reversion=mean_reversion, volatility=volatility,
      initial_discount_rate_fn=reference_rate_fn, corr_matrix=corr_matrix, dtype=dt_)
  batch_shape = expiries.shape
  m = pay_t.shape[-1]
  exp_rep = np.repeat(expiries[..., None], m, axis=-1)                  # swaption_pricing.py:292
  exp_flat = exp_rep.reshape(-1, m)[:, 0]
  pay_flat = np.broadcast_to(pay_t, batch_shape + (m,)).reshape(-1, m)
  dcf_flat = np.broadcast_to(dcf, batch_shape + (m,)).reshape(-1, m)
  cpn_flat = np.broadcast_to(coupon, batch_shape + (m,)).reshape(-1, m)
  ntl_flat = np.broadcast_to(ntl, batch_shape).reshape(-1)
  payer_flat = np.broadcast_to(payer, batch_shape).reshape(-1)
  # swaption_util.py:94-98: the simulation times are the sorted expiries (one per payment)
  sim_times = (np.sort(exp_rep.reshape(-1)) if times is None
               else _tensor.to_numpy(times, dt_).reshape(-1))
  plan, _, entry_of, inverse, y_entries, sim_times = model._plan(
      sim_times, time_step, num_time_steps, num_samples, random_type, seed, skip)
  try:
    sim_idx = np.searchsorted(sim_times, exp_flat, side='left')         # swaption_util.py:139
    descs = []
    for b in range(exp_flat.shape[0]):
      entry = int(entry_of[inverse[sim_idx[b]]])
      if entry == 0:
        raise ValueError('expiries must be positive.')
      y_e = model._y_at(sim_times[sim_idx[b]:sim_idx[b] + 1], y_entries[entry][None])[0]
      descs.append(_RawPayoff(_swaption_desc(
          model, entry, y_e, exp_flat[b], pay_flat[b], cpn_flat[b], dcf_flat[b], payer_flat[b],
          ntl_flat[b])))
    out = []
    for c0 in range(0, len(descs), _lib.MAX_PAYOFFS):
      out.append(distributed.price_sums_host(plan, descs[c0:c0 + _lib.MAX_PAYOFFS]))
    sums = np.concatenate(out, axis=0)
  finally:
    plan.close()
  n = float(plan.num_samples)
  price_ = (sums[:, 0] / n).astype(dt_).reshape(batch_shape)
  if not return_stats:
    return price_
  var = np.maximum(sums[:, 1] / n - (sums[:, 0] / n)**2, 0.0)
  return price_, np.sqrt(var / n).reshape(batch_shape), sums[:, 2].reshape(batch_shape)
