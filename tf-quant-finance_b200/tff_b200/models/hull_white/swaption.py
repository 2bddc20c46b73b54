"""Monte-Carlo European swaption pricing under the one-factor Hull-White model.

Drop-in for the simulation branch of
`tf_quant_finance.models.hull_white.swaption_price`
(`models/hull_white/swaption.py:41-312`) together with
`discount_factors_and_bond_prices_from_samples`
(`models/hjm/swaption_util.py:28-170`).

The reference materialises the short-rate paths `[N, k, 1]`, the bond-price
tensor `[N, m, k, 1]` (of which one time column is used) and a `[k, k]`
cumulative-sum matmul per path.  Here the fused kernel carries the path
integral of the short rate next to the OU state and evaluates
  payoff = notional * max(+-DF(t_e) (1 - P_N - sum_j c_j tau_j P_j), 0)
in registers when the path reaches the expiry step; nothing is stored.
"""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import distributed
from tff_b200.models import utils
from tff_b200.models.hull_white import _exact
from tff_b200.models.hull_white import one_factor


def _swaption_desc(model, step, expiry, pay_times, coupon, dcf, is_payer,
                   notional):
  """tqf_payoff_desc of one swaption evaluated after `step` steps."""
  dt_ = model._dtype
  k = model._tables.k
  y = model._tables.y_t(np.asarray([expiry], dtype=dt_))[0]
  rate = lambda t: _exact.discount_rate(model._initial_discount_rate_fn, t, dt_)
  t_e = np.asarray(expiry, dtype=dt_)
  ln_p0_ratio = -(rate(pay_times) * pay_times) + rate(t_e) * t_e
  g = (1. - np.exp(-k * (pay_times - t_e))) / k
  d = _lib.PayoffDesc()
  d.kind = _lib.PAYOFF_HW_SWAPTION
  d.expiry_step = int(step)
  d.num_payments = int(pay_times.shape[0])
  d.is_payer = int(bool(is_payer))
  d.scale = float(notional)
  coef = np.array(coupon * dcf, dtype=np.float64)
  coef[-1] += 1.0                      # float leg: 1 - P(t_e, T_N)
  for j in range(pay_times.shape[0]):
    d.pay_g[j] = float(g[j])
    d.pay_k[j] = float(ln_p0_ratio[j] - 0.5 * y * g[j]**2)
    d.pay_coef[j] = float(coef[j])
  return d


class _RawPayoff:
  def __init__(self, d):
    self._d = d

  def desc(self):
    return self._d


class _GridPricing:
  """The fused HW1F pricing plan of a batch of claims over one simulation grid:
  `plan` (device tables), `payoffs` (one descriptor per claim, evaluated when the
  path reaches the claim's expiry step), `num_steps`.  `sums()` runs the fused
  kernel and returns the payoff sums `[len(payoffs), 4]` (host); it may be called
  repeatedly (benchmarks re-price on the same plan)."""

  def __init__(self, plan, payoffs, num_steps):
    self.plan, self.payoffs, self.num_steps = plan, payoffs, num_steps
    self.num_samples = float(plan.num_samples)

  def sums_dev(self, chunk=0):
    c0 = chunk * _lib.MAX_PAYOFFS
    return distributed.price_sums(self.plan, self.payoffs[c0:c0 + _lib.MAX_PAYOFFS])

  def sums(self):
    out = []
    for c in range((len(self.payoffs) + _lib.MAX_PAYOFFS - 1) // _lib.MAX_PAYOFFS):
      c0 = c * _lib.MAX_PAYOFFS
      out.append(distributed.price_sums_host(self.plan, self.payoffs[c0:c0 + _lib.MAX_PAYOFFS]))
    return np.concatenate(out, axis=0)

  def close(self):
    self.plan.close()


def _plan_on_grid(model, sim_times, expiries, make_desc, num_samples,
                  random_type, seed, skip):
  """Builds the fused HW1F pricing plan over `sim_times` (`_GridPricing`).

  `make_desc(b, step)` builds the payoff descriptor of claim `b`, which is
  evaluated when its path reaches simulation step `step` (the step that lands
  on `expiries[b]`).  The path discount factor follows the reference's
  convention DF(t_j) = exp(-sum_{i<=j} r(t_i) dt_i) with dt_0 = 0
  (`hjm/swaption_util.py:126`, `hjm/zero_coupon_bond_option_util.py:102-113`).
  """
  dt_ = model._dtype

  def integral_weights(all_times, idx):
    # the step that lands on sim time j >= 1 carries t_j - t_{j-1}
    w = np.zeros(all_times.shape[0] - 1, dtype=dt_)
    dts = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]]).astype(dt_)
    for j, i in enumerate(idx):
      if i >= 1:
        w[i - 1] += dts[j]
    return w

  plan, _, all_times, idx = model._exact_plan(
      sim_times, int(num_samples), random_type, seed, skip, None, None,
      integral_weights_fn=integral_weights)
  try:
    sim_idx = np.searchsorted(sim_times, expiries, side='left')
    descs = []
    for b in range(expiries.shape[0]):
      step = int(idx[sim_idx[b]])
      if step == 0:
        # expiry at t = 0: the grid is [0, 0, ...] and the zero-length first
        # step leaves the state at its initial value (x = 0, integral = 0).
        if not (all_times.shape[0] > 1 and all_times[1] == all_times[0]):
          raise ValueError('expiries must be non-negative.')
        step = 1
      descs.append(_RawPayoff(make_desc(b, step)))
  except Exception:
    plan.close()
    raise
  return _GridPricing(plan, descs, plan.num_steps)


def _price_on_grid(model, sim_times, expiries, make_desc, num_samples,
                   random_type, seed, skip):
  """Payoff sums `[len(expiries), 4]` of the fused HW1F price kernel over
  `sim_times` and the number of simulated paths (see `_plan_on_grid`)."""
  gp = _plan_on_grid(model, sim_times, expiries, make_desc, num_samples,
                     random_type, seed, skip)
  try:
    sums = gp.sums()
  finally:
    gp.close()
  return sums, gp.num_samples


def swaption_price(*,
                   expiries,
                   floating_leg_start_times,
                   floating_leg_end_times,
                   fixed_leg_payment_times,
                   floating_leg_daycount_fractions,
                   fixed_leg_daycount_fractions,
                   fixed_leg_coupon,
                   reference_rate_fn,
                   mean_reversion,
                   volatility,
                   notional=None,
                   is_payer_swaption=True,
                   use_analytic_pricing=True,
                   num_samples=100,
                   random_type=None,
                   seed=None,
                   skip=0,
                   time_step=None,
                   dtype=None,
                   name=None,
                   return_stats=False,
                   _plan_only=False):
  """European swaption prices of shape `expiries.shape` (numpy float array).

  Same arguments as the reference.  `use_analytic_pricing=True` (the default, as in
  the reference) evaluates the Jamshidian closed form on the host;
  `use_analytic_pricing=False` runs the fused Monte-Carlo kernel.
  """
  del floating_leg_daycount_fractions, floating_leg_start_times
  del floating_leg_end_times, name
  dt_ = _tensor.infer_dtype(expiries, dtype, default=np.float32)
  expiries = _tensor.to_numpy(expiries, dt_)
  pay_t = _tensor.to_numpy(fixed_leg_payment_times, dt_)
  dcf = np.broadcast_to(_tensor.to_numpy(fixed_leg_daycount_fractions, dt_), pay_t.shape)
  coupon = np.broadcast_to(_tensor.to_numpy(fixed_leg_coupon, dt_), pay_t.shape)
  if expiries.ndim < pay_t.ndim - 1:
    raise ValueError('Swaption expiries not specified for all swaptions '
                     'in the batch. Expected rank {} but received {}.'.format(
                         pay_t.ndim - 1, expiries.ndim))
  notional = np.asarray(1.0 if notional is None else _tensor.to_numpy(notional, dt_), dtype=dt_)
  is_payer = np.asarray(_tensor.to_numpy(is_payer_swaption), dtype=bool)
  model = one_factor.HullWhiteModel1F(mean_reversion, volatility,
                                      reference_rate_fn, dtype=dt_)
  if use_analytic_pricing:
    # Jamshidian decomposition: a closed form evaluated on the host (swaption.py:726-814)
    if model._tables is None:
      raise ValueError('The paramerization of `mean_reversion` and/or `volatility` does not '
                       'support analytic computation of bond option variance.')
    from tff_b200.models.hull_white import _analytic  # pylint: disable=g-import-not-at-top
    price = _analytic.swaption_price(model, expiries, pay_t, dcf, coupon, notional, is_payer)
    return price.astype(dt_)
  if time_step is None:
    raise ValueError('`time_step` must be provided for simulation '
                     'based bond option valuation.')
  if model._tables is None:
    raise NotImplementedError(
        'swaption_price needs constant mean reversion and constant or '
        'piecewise-constant volatility (exact discretisation).')
  batch_shape = expiries.shape
  m = pay_t.shape[-1]
  exp_flat = np.broadcast_to(expiries[..., None], batch_shape + (m,)).reshape(-1, m)[:, 0]
  pay_flat = np.broadcast_to(pay_t, batch_shape + (m,)).reshape(-1, m)
  dcf_flat = np.broadcast_to(dcf, batch_shape + (m,)).reshape(-1, m)
  cpn_flat = np.broadcast_to(coupon, batch_shape + (m,)).reshape(-1, m)
  ntl_flat = np.broadcast_to(notional, batch_shape).reshape(-1)
  payer_flat = np.broadcast_to(is_payer, batch_shape).reshape(-1)

  # sim_times: unique expiries plus the uniform grid (swaption.py:284-288)
  sim_times = np.unique(exp_flat)
  longest = sim_times.max()
  sim_times = np.sort(np.concatenate(
      [sim_times, utils._tf_range(time_step, longest, time_step, dt_)]),
                      kind='stable').astype(dt_)

  def make_desc(b, step):
    return _swaption_desc(model, step, exp_flat[b], pay_flat[b], cpn_flat[b],
                          dcf_flat[b], payer_flat[b], ntl_flat[b])
  if _plan_only:
    return _plan_on_grid(model, sim_times, exp_flat, make_desc, num_samples,
                         random_type, seed, skip)
  sums, n = _price_on_grid(model, sim_times, exp_flat, make_desc, num_samples,
                           random_type, seed, skip)
  price = (sums[:, 0] / n).astype(dt_).reshape(batch_shape)
  if not return_stats:
    return price
  var = np.maximum(sums[:, 1] / n - (sums[:, 0] / n)**2, 0.0)
  return price, np.sqrt(var / n).reshape(batch_shape), sums[:, 2].reshape(batch_shape)


def _unique_in_order(a):
  """`tf.unique`: distinct values in order of first appearance, and the inverse."""
  _, first, inv = np.unique(a, return_index=True, return_inverse=True)
  order = np.argsort(first, kind='stable')
  rank = np.empty_like(order)
  rank[order] = np.arange(order.shape[0])
  return a[np.sort(first)], rank[inv]


def bermudan_swaption_price(*,
                            exercise_times,
                            floating_leg_start_times,
                            floating_leg_end_times,
                            fixed_leg_payment_times,
                            floating_leg_daycount_fractions,
                            fixed_leg_daycount_fractions,
                            fixed_leg_coupon,
                            reference_rate_fn,
                            mean_reversion,
                            volatility,
                            notional=None,
                            is_payer_swaption=True,
                            use_finite_difference=False,
                            lsm_basis=None,
                            num_samples=100,
                            random_type=None,
                            seed=None,
                            skip=0,
                            time_step=None,
                            time_step_finite_difference=None,
                            num_grid_points_finite_difference=101,
                            dtype=None,
                            name=None):
  """Bermudan swaption prices by Longstaff-Schwartz on Hull-White short-rate
  paths (`models/hull_white/swaption.py:314-724`, Monte-Carlo branch).

  `exercise_times`: `batch_shape + [E]`; the leg arrays: `batch_shape + [E, m]`
  (for every exercise date the remaining payments, padded).  Returns a numpy
  array of shape `batch_shape`.

  The short-rate state and its path integral come from the fused HW1F kernel
  (nothing but the `[N, k, 2]` state is stored), the exercise values
  `+-(1 - P_N - sum_j c_j tau_j P_j)` are tabulated on the device from the
  closed-form bond prices, and the regression / exercise passes are the LSM
  kernels with one discount curve per path.  As in the reference the swaption
  is valued as a payer (`swaption.py:562`); the finite-difference branch is a
  PDE solver outside the Monte-Carlo hot path and is not provided.
  """
  del floating_leg_daycount_fractions, floating_leg_start_times, floating_leg_end_times
  del name, time_step_finite_difference, num_grid_points_finite_difference, is_payer_swaption
  import torch
  from tff_b200.models.longstaff_schwartz import lsm
  from tff_b200.models.longstaff_schwartz import payoff_utils
  dt_ = _tensor.infer_dtype(exercise_times, dtype, default=np.float32)
  ex = _tensor.to_numpy(exercise_times, dt_)
  pay_t = _tensor.to_numpy(fixed_leg_payment_times, dt_)
  dcf = np.broadcast_to(_tensor.to_numpy(fixed_leg_daycount_fractions, dt_), pay_t.shape)
  coupon = np.broadcast_to(_tensor.to_numpy(fixed_leg_coupon, dt_), pay_t.shape)
  ntl = np.asarray(1.0 if notional is None else _tensor.to_numpy(notional, dt_), dtype=dt_)
  if use_finite_difference:
    raise NotImplementedError(
        'The finite-difference Bermudan valuation is a PDE solver outside the '
        'B200 Monte-Carlo hot path; call with use_finite_difference=False.')
  if ex.ndim < pay_t.ndim - 1:
    raise ValueError('Swaption exercise times not specified for all '
                     'swaptions in the batch. Expected rank '
                     '{} but received {}.'.format(pay_t.ndim - 1, ex.ndim))
  if time_step is None:
    raise ValueError('`time_step` must be provided for LSM valuation.')
  basis_fn = lsm.make_polynomial_basis(2) if lsm_basis is None else lsm_basis
  model = one_factor.HullWhiteModel1F(mean_reversion, volatility,
                                      reference_rate_fn, dtype=dt_)
  if model._tables is None:
    raise NotImplementedError(
        'bermudan_swaption_price needs constant mean reversion and constant or '
        'piecewise-constant volatility (exact discretisation).')
  batch_shape = ex.shape[:-1]
  n_ex, m = ex.shape[-1], pay_t.shape[-1]
  nb = int(np.prod(batch_shape)) if batch_shape else 1
  ex_flat = ex.reshape(nb, n_ex)
  pay_flat = np.broadcast_to(pay_t, batch_shape + (n_ex, m)).reshape(nb, n_ex, m)
  coef = (np.broadcast_to(coupon, batch_shape + (n_ex, m)) *
          np.broadcast_to(dcf, batch_shape + (n_ex, m))).reshape(nb, n_ex, m).astype(np.float64)
  coef[..., -1] += 1.0                           # float leg: 1 - P(t_e, T_N)

  # unique exercise dates in order of first appearance; simulation grid
  # (swaption.py:571-574, 611-615; `longest` is the LAST unique date)
  uniq, ex_index = _unique_in_order(ex_flat.reshape(-1))
  ex_index = ex_index.reshape(nb, n_ex)
  longest = uniq[-1]
  sim_times = np.unique(np.concatenate(
      [uniq, utils._tf_range(time_step, longest, time_step, dt_)])).astype(dt_)
  k_sim = sim_times.shape[0]

  def integral_weights(all_times, idx):
    w = np.zeros(all_times.shape[0] - 1, dtype=dt_)
    dts = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]]).astype(dt_)
    for j, i in enumerate(idx):
      if i >= 1:
        w[i - 1] += dts[j]
    return w
  plan, record_slot, _, _ = model._exact_plan(
      sim_times, int(num_samples), random_type, seed, skip, None, None,
      integral_weights_fn=integral_weights)
  try:
    state = plan.paths(record_slot, k_sim)                   # [N, k, 2] = (x, integral)
  finally:
    plan.close()
  n = int(state.shape[0])
  dev, td = state.device, state.dtype
  sim_idx = np.searchsorted(sim_times, uniq, side='left')   # [U]
  sel = torch.as_tensor(sim_idx, device=dev)
  x_u = state[:, sel, 0]                                      # [N, U]
  f0 = torch.as_tensor(model._fwd(uniq), device=dev, dtype=td)
  short_rate = (x_u + f0[None, :]).unsqueeze(-1).contiguous()      # [N, U, 1]
  df_u = torch.exp(-state[:, sel, 1]).unsqueeze(1).contiguous()    # [N, 1, U]

  # bond prices at the exercise dates: P(t, T_j) = exp(kk_j - g_j x(t))
  kconst = model._tables.k
  t_e = np.repeat(ex_flat[..., None], m, axis=-1)             # [nb, E, m]
  rate = lambda t: _exact.discount_rate(model._initial_discount_rate_fn, t, dt_)
  g = (1. - np.exp(-kconst * (pay_flat - t_e))) / kconst
  y = model._tables.y_t(t_e.reshape(-1)).reshape(t_e.shape)
  kk = -(rate(pay_flat) * pay_flat) + rate(t_e) * t_e - 0.5 * y * g**2
  u_count = uniq.shape[0]
  # exercise values of every (swaption, exercise date) on every path: one kernel
  # (tqf_hw_exercise_values) instead of a gather / sum / scatter per pair
  g_d = torch.as_tensor(np.ascontiguousarray(g, dtype=np.float64), device=dev)
  k_d = torch.as_tensor(np.ascontiguousarray(kk, dtype=np.float64), device=dev)
  c_d = torch.as_tensor(np.ascontiguousarray(coef, dtype=np.float64), device=dev)
  slot_d = torch.as_tensor(np.ascontiguousarray(ex_index, dtype=np.int32), device=dev)
  values = torch.zeros((u_count, n, nb), device=dev, dtype=td)
  x_u = x_u.contiguous()                                      # [N, U]
  _lib.require_cuda()
  _lib.check(_lib.lib().tqf_hw_exercise_values(
      x_u.data_ptr(), x_u.stride(0), x_u.stride(1), g_d.data_ptr(), k_d.data_ptr(),
      c_d.data_ptr(), slot_d.data_ptr(), n, nb, n_ex, m, _tensor.tqf_dtype(dt_),
      values.data_ptr(), _tensor.current_stream_ptr()))
  price = lsm.least_square_mc(
      short_rate, np.arange(u_count), payoff_utils.make_tabulated_payoff(values),
      basis_fn, discount_factors=df_u, dtype=dt_)
  price = (np.broadcast_to(ntl, batch_shape).reshape(-1) * price).astype(dt_)
  return price.reshape(batch_shape)
