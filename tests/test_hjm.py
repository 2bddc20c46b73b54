"""Gaussian / quasi-Gaussian HJM and the Monte-Carlo HJM swaption, bond-option and cap / floor
prices (`models/hjm/{quasi_gaussian_hjm,gaussian_hjm,swaption_pricing,swaption_util,
zero_coupon_bond_option,zero_coupon_bond_option_util,cap_floor}.py`).

CPU: the oracle against the reference's own values (`gaussian_hjm_test.py:224-283`,
`swaption_pricing_test.py:46-165, 321-356`), and the product's HOST tables (grid,
deterministic y, per-step coefficients, payoff descriptors) replayed in numpy on the
oracle's draws against the oracle's state-space simulation.  GPU: the kernels against
the oracle on the same seeds."""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import hjm as ohjm
from oracle import models as omodels

RATE = lambda t: 0.01 + 0 * t
SWAPTION = dict(expiries=np.array([1.0]), fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
                fixed_leg_daycount_fractions=0.25 * np.ones(4),
                fixed_leg_coupon=0.011 * np.ones(4), reference_rate_fn=RATE, notional=100.,
                seed=[1, 2], dtype=np.float64)


def test_oracle_bond_price_kats():
  # gaussian_hjm_test.py:224-283 (the test hands f(0,t) = 0.01 in as the discount rate)
  for dim, expected in ((1, [0.9803327113840525, 0.9803218405347454, 0.9803116028646381]),
                        (2, [0.9707109604475661, 0.9706894322583266, 0.9706691582097785])):
    p = ohjm.GaussianHJM(dim, [0.03] * dim, [0.005] * dim, RATE)
    t = np.array([1.0, 2.0, 3.0])
    got = p.discount_bond_price(0.01 * np.ones((3, dim)), t, t + 1.0)
    np.testing.assert_allclose(got, expected, rtol=1e-12)


def test_oracle_swaption_reference_values():
  rt = odraws.RandomType.STATELESS_ANTITHETIC
  kw = dict(SWAPTION, num_samples=50_000, random_type=rt)
  one = dict(num_hjm_factors=1, mean_reversion=[0.03], volatility=[0.02])
  # swaption_pricing_test.py:46-90 (both grids), 92-127 (receiver), tolerance 1e-2 there
  assert abs(ohjm.swaption_price_mc(time_step=0.1, **one, **kw)[0] - 0.7163243383624043) < 1e-2
  assert abs(ohjm.swaption_price_mc(num_time_steps=11, **one, **kw)[0] - 0.7163243383624043) < 1e-2
  assert abs(ohjm.swaption_price_mc(time_step=0.1, is_payer_swaption=False, **one, **kw)[0]
             - 0.813482544626056) < 1e-2
  # :129-165 time-dependent volatility through a callable of (t, r), tolerance 1e-3
  pw = omodels.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  got = ohjm.swaption_price_mc(num_hjm_factors=1, mean_reversion=[0.03], time_step=0.1,
                               volatility=lambda t, r: pw(np.asarray([t])), **kw)
  assert abs(got[0] - 0.5593057004094042) < 1e-3
  # :321-356 two factors
  got = ohjm.swaption_price_mc(num_hjm_factors=2, mean_reversion=[0.03, 0.06],
                               volatility=[0.02, 0.01], time_step=0.1,
                               **dict(kw, num_samples=25_000))
  assert abs(got[0] - 0.802226) < 1e-2


def _replay(model, times, num_samples, random_type, seed, skip=0, time_step=None,
            num_time_steps=None, integral_weights_fn=None):
  """numpy twin of HjmModel::step (csrc/tqf_paths_kernel.cuh) over the product's host
  tables with the oracle's draws: (state at the requested times [N, k, F + 1], y)."""
  dt_ = model._dtype
  times = np.asarray(times, dtype=dt_)
  f = model._factors
  grid, idx, all_times, keep_mask = model._grids(times, time_step, num_time_steps)
  from tff_b200 import engine
  num_steps, grid_slot = engine.record_plan(keep_mask, grid.shape[0])
  weights = None
  if integral_weights_fn is not None:
    entries = [max(e for e, g in enumerate(grid_slot) if g == i) for i in idx]
    weights = integral_weights_fn(all_times, entries)
  table, y_entries = model._tables(all_times, weights)
  nfs = model._draws_per_step()
  z = odraws.generate_mc_normal_draws(nfs, all_times.shape[0] - 1, num_samples, random_type,
                                      seed=seed, skip=skip, dtype=dt_)           # [S, N, nfs]
  x = np.zeros((num_samples, f), dtype=dt_)
  integ = np.zeros(num_samples, dtype=dt_)
  slots = {0: (x.copy(), integ.copy())} if grid_slot[0] >= 0 else {}
  for s in range(num_steps):
    c = table[s].astype(dt_)
    dw = z[s][:, :f] * c[1]
    a0, k = c[2:2 + f], c[2 + f:2 + 2 * f]
    b = c[2 + 2 * f:2 + 2 * f + f * f].reshape(f, f)
    xn = (x + c[0] * (a0 - k * x)) + dw @ b.T
    integ = integ + (c[2 + 2 * f + f * f] * x.sum(-1) + c[3 + 2 * f + f * f] * xn.sum(-1)
                     + c[4 + 2 * f + f * f])
    x = xn
    if grid_slot[s + 1] >= 0:
      slots[int(grid_slot[s + 1])] = (x.copy(), integ.copy())
  state = np.stack([np.concatenate([slots[int(g)][0], slots[int(g)][1][:, None]], -1)
                    for g in idx], 1)
  return state, times


@pytest.mark.parametrize('factors,grid', [(1, dict(time_step=0.1)), (1, dict(num_time_steps=11)),
                                          (2, dict(time_step=0.07))])
def test_host_tables_replay_matches_oracle_quasi_gaussian(factors, grid):
  from tff_b200.models import hjm
  mr = [0.03, 0.06][:factors]
  vol = [0.02, 0.01][:factors]
  corr = None if factors == 1 else [[1.0, 0.4], [0.4, 1.0]]
  times = np.array([0.3, 1.0, 1.7, 1.7])                    # a duplicate, as the swaption grid has
  rt = odraws.RandomType.STATELESS_ANTITHETIC
  model = hjm.QuasiGaussianHJM(factors, mr, vol, RATE, corr_matrix=corr, dtype=np.float64)
  state, _ = _replay(model, times, 64, rt, [4, 2], **grid)
  want = ohjm.QuasiGaussianHJM(factors, mr, vol, RATE, corr_matrix=corr)
  rate, df, x, y = want.sample_paths(times, 64, random_type=rt, seed=[4, 2], **grid)
  np.testing.assert_allclose(state[..., :factors], x, rtol=1e-12, atol=1e-16)
  np.testing.assert_allclose(np.exp(-state[..., factors]), df, rtol=1e-13)
  np.testing.assert_allclose(state[..., :factors].sum(-1) + 0.01, rate, rtol=1e-12, atol=1e-16)


def test_host_tables_replay_matches_oracle_gaussian():
  from tff_b200.math import piecewise
  from tff_b200.models import hjm
  rt = odraws.RandomType.STATELESS
  times = np.array([0.1, 0.5, 1.0, 2.0])
  for factors, vol_o, vol_p, grid in (
      (1, [0.01], [0.01], dict(num_time_steps=21)),
      (2, omodels.PiecewiseConstantFunc([[0.5, 1.0], [0.5, 1.0]], [[0.005, 0.008, 0.005]] * 2,
                                        dtype=np.float64),
       piecewise.PiecewiseConstantFunc([[0.5, 1.0], [0.5, 1.0]], [[0.005, 0.008, 0.005]] * 2,
                                       dtype=np.float64), dict(time_step=0.1))):
    mr = [0.03, 0.1][:factors]
    corr = None if factors == 1 else [[1.0, 0.5], [0.5, 1.0]]
    model = hjm.GaussianHJM(factors, mr, vol_p, RATE, corr_matrix=corr, dtype=np.float64)
    want = ohjm.GaussianHJM(factors, mr, vol_o, RATE, corr_matrix=corr)
    np.testing.assert_allclose(model.state_y(times), want.state_y(times), rtol=1e-12, atol=1e-20)
    state, _ = _replay(model, times, 32, rt, [7, 9], **grid)
    rate, df, x, _ = want.sample_paths(times, 32, random_type=rt, seed=[7, 9], **grid)
    np.testing.assert_allclose(state[..., :factors], x, rtol=1e-11, atol=1e-16)
    np.testing.assert_allclose(np.exp(-state[..., factors]), df, rtol=1e-13)
  # bond price KAT through the product's closed form (gaussian_hjm_test.py:224-283)
  p = hjm.GaussianHJM(2, [0.03, 0.03], [0.005, 0.005], RATE, dtype=np.float64)
  t = np.array([1.0, 2.0, 3.0])
  np.testing.assert_allclose(
      p.discount_bond_price(0.01 * np.ones((3, 2)), t, t + 1.0),
      [0.9707109604475661, 0.9706894322583266, 0.9706691582097785], rtol=1e-12)


@pytest.mark.parametrize('factors', [1, 2])
def test_swaption_descriptor_replay_matches_oracle_price(factors):
  """The payoff descriptor the pricer hands to the kernel, evaluated in numpy on the replayed
  state exactly as the kernel's swaption branch does, against the oracle's price."""
  from tff_b200.models import hjm
  from tff_b200.models.hjm import swaption_pricing as sp
  mr, vol = [0.03, 0.06][:factors], [0.02, 0.01][:factors]
  rt = odraws.RandomType.STATELESS_ANTITHETIC
  n = 4000
  model = hjm.QuasiGaussianHJM(factors, mr, vol, RATE, dtype=np.float64)
  sim_times = np.array([1.0, 1.0, 1.0, 1.0])
  state, _ = _replay(model, sim_times, n, rt, [1, 2], time_step=0.1)
  grid, idx, all_times, keep_mask = model._grids(sim_times, 0.1, None)
  _, y_entries = model._tables(all_times)
  from tff_b200 import engine
  _, grid_slot = engine.record_plan(keep_mask, grid.shape[0])
  entry = max(e for e, g in enumerate(grid_slot) if g == idx[0])
  pay = SWAPTION['fixed_leg_payment_times']
  d = sp._swaption_desc(model, entry, y_entries[entry], 1.0, pay, 0.011 * np.ones(4),
                        0.25 * np.ones(4), True, 100.0)
  x, integ = state[:, 0, :factors], state[:, 0, factors]
  acc = np.zeros(n)
  for j in range(d.num_payments):
    e = d.pay_k[j] - sum(d.pay_g[j * factors + i] * x[:, i] for i in range(factors))
    acc += d.pay_coef[j] * np.exp(e)
  price = d.scale * np.maximum(np.exp(-integ) * (1.0 - acc), 0.0).mean()
  want = ohjm.swaption_price_mc(num_hjm_factors=factors, mean_reversion=mr, volatility=vol,
                                time_step=0.1, num_samples=n, random_type=rt, **SWAPTION)
  np.testing.assert_allclose(price, want[0], rtol=1e-11)


def test_oracle_bond_option_and_cap_reference_values():
  rt = odraws.RandomType.STATELESS_ANTITHETIC
  # zero_coupon_bond_option_test.py:37-61 (tolerance 1e-2 there), 63-96 (time-dependent vol)
  exp, mat = np.array([1.0]), np.array([5.0])
  strikes = np.exp(-0.01 * mat) / np.exp(-0.01 * exp)
  kw = dict(strikes=strikes, expiries=exp, maturities=mat, discount_rate_fn=RATE, time_step=0.1,
            random_type=rt, seed=[1, 2])
  got = ohjm.bond_option_price_mc(dim=1, mean_reversion=[0.03], volatility=[0.02],
                                  num_samples=100_000, **kw)
  assert got.shape == (1,) and abs(got[0] - 0.02817777) < 1e-2
  # :183-209 two factors
  got = ohjm.bond_option_price_mc(dim=2, mean_reversion=[0.03, 0.06], volatility=[0.02, 0.01],
                                  num_samples=50_000, **kw)
  assert abs(got[0] - 0.03111126) < 1e-2
  # cap_floor_test.py:41-68: the analytic Hull-White value 0.4072088 (tolerance 1e-3 + 1e-3 rel
  # there; a single seed scatters by ~1e-3, the mean over seeds sits within 5e-4)
  cap = dict(strikes=0.01 * np.ones(4), expiries=np.array([0.0, 0.25, 0.5, 0.75]),
             maturities=np.array([0.25, 0.5, 0.75, 1.0]), daycount_fractions=0.25 * np.ones(4),
             notional=100.0, dim=1, mean_reversion=[0.03], volatility=[0.02], reference_rate_fn=RATE,
             num_samples=100_000, time_step=0.1, random_type=rt)
  prices = [ohjm.cap_floor_price_mc(seed=seed, **cap) for seed in ([42, 42], [1, 2], [3, 4], [5, 6])]
  assert abs(prices[0] - 0.4072088281493774) < 2.5e-3
  assert abs(np.mean(prices) - 0.4072088281493774) < 5e-4
  # :70-103 piecewise-constant volatility
  pw = omodels.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  got = ohjm.cap_floor_price_mc(seed=[42, 42], **dict(cap, volatility=lambda t, r: pw(np.asarray([t]))))
  assert abs(got - 0.2394242699989869) < 2.5e-3


def test_oracle_bond_option_batches_and_call_put_reference_values():
  # zero_coupon_bond_option_test.py:125-153 (2-d batch), 262-290 (mixed batch, two factors),
  # 292-320 (calls and puts); tolerance 1e-2 there
  rt = odraws.RandomType.STATELESS_ANTITHETIC
  kw = dict(discount_rate_fn=RATE, num_samples=50_000, time_step=0.1, random_type=rt, seed=[1, 2])
  exp = np.array([[1.0, 1.0], [2.0, 2.0]])
  mat = np.array([[5.0, 5.0], [4.0, 4.0]])
  got = ohjm.bond_option_price_mc(strikes=np.exp(-0.01 * mat) / np.exp(-0.01 * exp), expiries=exp,
                                  maturities=mat, dim=1, mean_reversion=[0.03], volatility=[0.02], **kw)
  assert got.shape == (2, 2)
  np.testing.assert_allclose(got, [[0.02817777, 0.02817777], [0.02042677, 0.02042677]], rtol=1e-2,
                             atol=1e-2)
  exp, mat = np.array([1.0, 1.0, 2.0]), np.array([5.0, 6.0, 4.0])
  two = dict(dim=2, mean_reversion=[0.03, 0.06], volatility=[0.02, 0.01])
  got = ohjm.bond_option_price_mc(strikes=np.exp(-0.01 * mat) / np.exp(-0.01 * exp), expiries=exp,
                                  maturities=mat, **two, **kw)
  np.testing.assert_allclose(got, [0.03115176, 0.03789011, 0.02266191], rtol=1e-2, atol=1e-2)
  got = ohjm.bond_option_price_mc(strikes=np.exp(-0.01 * mat) / np.exp(-0.01 * exp) - 0.01, expiries=exp,
                                  maturities=mat, is_call_options=[True, False, False], **two, **kw)
  np.testing.assert_allclose(got, [0.03620415, 0.03279728, 0.01784987], rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('factors', [1, 2])
def test_bond_option_descriptor_replay_matches_oracle_price(factors):
  """The bond-option route on the host: simulation grid, the reference's discounting weights
  (dt_0 = 0 over the SIMULATION times) and the one-payment payoff descriptors, evaluated in numpy
  on the replayed state exactly as the kernel does, against the oracle's price."""
  from tff_b200.models import hjm
  from tff_b200.models.hjm import zero_coupon_bond_option as zcb
  from tff_b200 import engine
  mr, vol = [0.03, 0.06][:factors], [0.02, 0.01][:factors]
  rt = odraws.RandomType.STATELESS_ANTITHETIC
  n = 4000
  strikes = np.array([0.96, 0.97, 0.99, 0.95])
  expiries = np.array([1.0, 0.55, 0.25, 1.0])
  maturities = np.array([5.0, 2.0, 0.5, 3.0])
  is_call = np.array([True, False, True, False])
  model = hjm.QuasiGaussianHJM(factors, mr, vol, RATE, dtype=np.float64)
  sim_times, wfn = zcb._simulation_grid(expiries, np.float64(0.1), np.dtype(np.float64))
  state, _ = _replay(model, sim_times, n, rt, [1, 2], time_step=np.float64(0.1), integral_weights_fn=wfn)
  grid, idx, all_times, keep_mask = model._grids(sim_times, np.float64(0.1), None)
  _, grid_slot = engine.record_plan(keep_mask, grid.shape[0])
  _, y_entries = model._tables(all_times)
  got = []
  for b in range(4):
    j = int(np.searchsorted(sim_times, expiries[b]))
    entry = max(e for e, g in enumerate(grid_slot) if g == idx[j])
    d = zcb._bond_option_desc(model, entry, y_entries[entry], expiries[b], maturities[b], strikes[b],
                              is_call[b])
    x, integ = state[:, j, :factors], state[:, j, factors]
    p = d.pay_coef[0] * np.exp(d.pay_k[0] - sum(d.pay_g[i] * x[:, i] for i in range(factors)))
    sign = 1.0 if d.is_payer else -1.0
    got.append(d.scale * np.maximum(sign * np.exp(-integ) * (1.0 - p), 0.0).mean())
  want = ohjm.bond_option_price_mc(strikes=strikes, expiries=expiries, maturities=maturities,
                                   discount_rate_fn=RATE, dim=factors, mean_reversion=mr,
                                   volatility=vol, is_call_options=is_call, num_samples=n,
                                   random_type=rt, seed=[1, 2], time_step=0.1)
  np.testing.assert_allclose(got, want, rtol=1e-10)


def test_interior_duplicate_times_with_num_time_steps_are_refused():
  # `_grid_from_num_times` keeps duplicate request times; the Euler loop then leaves the
  # trailing TensorArray slots unwritten and the reference gathers one of them (zeros) for
  # every time AFTER an interior duplicate.  The engine refuses instead of returning zeros.
  from tff_b200.models import hjm
  model = hjm.QuasiGaussianHJM(1, [0.03], [0.02], RATE, dtype=np.float64)
  want = ohjm.QuasiGaussianHJM(1, [0.03], [0.02], RATE)
  _, _, x, _ = want.sample_paths(np.array([0.3, 1.0, 1.0, 1.7]), 8, num_time_steps=11,
                                 random_type=odraws.RandomType.STATELESS, seed=[1, 2])
  assert np.all(x[:, 3] == 0.0) and np.all(x[:, 1] != 0.0)        # what the reference returns
  grid, idx, all_times, keep_mask = model._grids(np.array([0.3, 1.0, 1.0, 1.7]), None, 11)
  from tff_b200 import engine
  _, grid_slot = engine.record_plan(keep_mask, grid.shape[0])
  assert int(idx[3]) not in set(int(g) for g in grid_slot)


def test_rate_dependent_volatility_is_refused():
  from tff_b200.models import hjm
  model = hjm.QuasiGaussianHJM(1, [0.03], lambda t, r: 0.02 * (1 + abs(r)), RATE,
                               dtype=np.float64)
  with pytest.raises(NotImplementedError):
    model._tables(np.array([0.0, 0.1, 0.2]))


# ------------------------------------------------------------------ GPU -----
def _np(t):
  return t.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize('factors,grid', [(1, dict(time_step=0.1)), (1, dict(num_time_steps=11)),
                                          (2, dict(time_step=0.05))])
def test_gpu_quasi_gaussian_paths_and_curves_match_oracle(factors, grid):
  import tff_b200 as tff
  mr = [0.03, 0.06][:factors]
  vol = [0.02, 0.01][:factors]
  corr = None if factors == 1 else [[1.0, 0.4], [0.4, 1.0]]
  times = np.array([0.3, 1.0, 1.7, 1.7])
  curve_times = np.array([0.0, 0.5, 1.0, 5.0])
  n = 3000
  for rt_name in ('STATELESS_ANTITHETIC', 'SOBOL'):
    rt = getattr(tff.math.random.RandomType, rt_name)
    ort = getattr(odraws.RandomType, rt_name)
    model = tff.models.hjm.QuasiGaussianHJM(factors, mr, vol, RATE, corr_matrix=corr,
                                            dtype=np.float64)
    want = ohjm.QuasiGaussianHJM(factors, mr, vol, RATE, corr_matrix=corr)
    rate, df, x, y = model.sample_paths(times, n, random_type=rt, seed=[4, 2], skip=3, **grid)
    wr, wdf, wx, wy = want.sample_paths(times, n, random_type=ort, seed=[4, 2], skip=3, **grid)
    np.testing.assert_allclose(_np(x), wx, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(_np(rate), wr, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(_np(df), wdf, rtol=1e-12)
    np.testing.assert_allclose(_np(y), wy, rtol=1e-12, atol=1e-20)
    p, rate2, df2 = model.sample_discount_curve_paths(times, curve_times, n, random_type=rt,
                                                      seed=[4, 2], skip=3, **grid)
    wp, _, _ = want.sample_discount_curve_paths(times, curve_times, n, random_type=ort,
                                                seed=[4, 2], skip=3, **grid)
    assert tuple(p.shape) == (n, 4, 4)
    np.testing.assert_allclose(_np(p), wp, rtol=1e-12)
    np.testing.assert_allclose(_np(df2), wdf, rtol=1e-12)


@pytest.mark.gpu
def test_gpu_gaussian_hjm_matches_oracle():
  import tff_b200 as tff
  rt, ort = tff.math.random.RandomType.STATELESS_ANTITHETIC, odraws.RandomType.STATELESS_ANTITHETIC
  times = np.array([0.1, 0.5, 1.0, 2.0])
  for factors, mr, vol, corr, grid in ((1, [0.03], [0.01], None, dict(num_time_steps=21)),
                                       (2, [0.03, 0.1], [0.005, 0.012], [[1.0, 0.5], [0.5, 1.0]],
                                        dict(time_step=0.1)),
                                       (3, [0.03, 0.1, 0.2], [0.005, 0.012, 0.007], None,
                                        dict(time_step=0.1))):
    model = tff.models.hjm.GaussianHJM(factors, mr, vol, RATE, corr_matrix=corr, dtype=np.float64)
    want = ohjm.GaussianHJM(factors, mr, vol, RATE, corr_matrix=corr)
    rate, df, x, y = model.sample_paths(times, 2000, random_type=rt, seed=[1, 2], **grid)
    wr, wdf, wx, wy = want.sample_paths(times, 2000, random_type=ort, seed=[1, 2], **grid)
    np.testing.assert_allclose(_np(x), wx, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(_np(rate), wr, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(_np(df), wdf, rtol=1e-12)
    np.testing.assert_allclose(_np(y), wy, rtol=1e-12, atol=1e-20)
  # gaussian_hjm_test.py:58-163: E[discount factor] = P(0, t)
  model = tff.models.hjm.GaussianHJM(1, [0.03], [0.01], RATE, dtype=np.float64)
  _, df, _, _ = model.sample_paths(times, 100_000, time_step=0.1, random_type=rt, seed=[1, 2])
  np.testing.assert_allclose(_np(df).mean(0), np.exp(-0.01 * times), rtol=1e-3)


@pytest.mark.gpu
def test_gpu_hjm_swaption_matches_oracle_and_reference_values():
  import tff_b200 as tff
  rt, ort = tff.math.random.RandomType.STATELESS_ANTITHETIC, odraws.RandomType.STATELESS_ANTITHETIC
  price = tff.models.hjm.swaption_price
  one = dict(num_hjm_factors=1, mean_reversion=[0.03], volatility=[0.02])
  two = dict(num_hjm_factors=2, mean_reversion=[0.03, 0.06], volatility=[0.02, 0.01])
  cases = [(one, dict(time_step=0.1), 50_000, 0.7163243383624043, 1e-2),
           (one, dict(num_time_steps=11), 50_000, 0.7163243383624043, 1e-2),
           (dict(one, is_payer_swaption=False), dict(time_step=0.1), 50_000, 0.813482544626056, 1e-2),
           (two, dict(time_step=0.1), 25_000, 0.802226, 1e-2),
           (dict(two, corr_matrix=[[1.0, 0.5], [0.5, 1.0]]), dict(time_step=0.1), 20_000, None, None)]
  for model_kw, grid, n, ref, tol in cases:
    got = price(num_samples=n, random_type=rt, **SWAPTION, **model_kw, **grid)
    want = ohjm.swaption_price_mc(num_samples=n, random_type=ort, **SWAPTION, **model_kw, **grid)
    assert got.shape == (1,) and got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=1e-10)
    if ref is not None:
      assert abs(got[0] - ref) < tol
  # a batch with different expiries, notionals and payer flags in one fused launch
  kw = dict(expiries=np.array([1.0, 2.0, 1.0]),
            fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0], [2.25, 2.5, 2.75, 3.0],
                                              [1.25, 1.5, 1.75, 2.0]]),
            fixed_leg_daycount_fractions=0.25 * np.ones((3, 4)),
            fixed_leg_coupon=0.011 * np.ones((3, 4)), reference_rate_fn=RATE,
            notional=np.array([100., 50., 100.]), is_payer_swaption=np.array([True, True, False]),
            seed=[1, 2], dtype=np.float64, num_samples=8000, time_step=0.1)
  got = price(random_type=rt, **one, **kw)
  want = ohjm.swaption_price_mc(random_type=ort, **one, **kw)
  np.testing.assert_allclose(got, want, rtol=1e-10)
  # time-dependent volatility through a callable (swaption_pricing_test.py:129-165)
  from tff_b200.math import piecewise
  pw = piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  opw = omodels.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  got = price(num_hjm_factors=1, mean_reversion=[0.03], volatility=lambda t, r: pw([float(t)]),
              num_samples=50_000, random_type=rt, time_step=0.1, **SWAPTION)
  want = ohjm.swaption_price_mc(num_hjm_factors=1, mean_reversion=[0.03],
                                volatility=lambda t, r: opw(np.asarray([t])), num_samples=50_000,
                                random_type=ort, time_step=0.1, **SWAPTION)
  np.testing.assert_allclose(got, want, rtol=1e-10)
  assert abs(got[0] - 0.5593057004094042) < 1e-3


@pytest.mark.gpu
def test_gpu_hjm_bond_option_and_cap_floor_match_oracle_and_reference_values():
  import tff_b200 as tff
  rt, ort = tff.math.random.RandomType.STATELESS_ANTITHETIC, odraws.RandomType.STATELESS_ANTITHETIC
  one = dict(dim=1, mean_reversion=[0.03], volatility=[0.02])
  two = dict(dim=2, mean_reversion=[0.03, 0.06], volatility=[0.02, 0.01])
  exp, mat = np.array([1.0]), np.array([5.0])
  strikes = np.exp(-0.01 * mat) / np.exp(-0.01 * exp)
  # zero_coupon_bond_option_test.py:37-61, 183-209 (tolerance 1e-2 there)
  for model_kw, n, ref in ((one, 100_000, 0.02817777), (two, 50_000, 0.03111126),
                           (dict(two, corr_matrix=[[1.0, 0.5], [0.5, 1.0]]), 20_000, None)):
    kw = dict(strikes=strikes, expiries=exp, maturities=mat, discount_rate_fn=RATE, time_step=0.1,
              seed=[1, 2], num_samples=n, **model_kw)
    got = tff.models.hjm.bond_option_price(random_type=rt, dtype=np.float64, **kw)
    want = ohjm.bond_option_price_mc(random_type=ort, **kw)
    assert got.shape == (1,) and got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=1e-10)
    if ref is not None:
      assert abs(got[0] - ref) < 1e-2
  # a batch: calls and puts, several expiries (one off the uniform grid), a 2-d batch shape
  kw = dict(strikes=np.array([[0.96, 0.97], [0.99, 0.95]]), expiries=np.array([[1.0, 0.55], [0.25, 1.0]]),
            maturities=np.array([[5.0, 2.0], [0.5, 3.0]]), discount_rate_fn=RATE, time_step=0.1,
            is_call_options=np.array([[True, False], [True, False]]), seed=[4, 2], num_samples=20_000)
  got, stderr, bad = tff.models.hjm.bond_option_price(random_type=rt, dtype=np.float64,
                                                      return_stats=True, **one, **kw)
  want = ohjm.bond_option_price_mc(random_type=ort, **one, **kw)
  assert got.shape == (2, 2) and np.all(bad == 0) and np.all(stderr > 0)
  np.testing.assert_allclose(got, want, rtol=1e-10)
  # cap_floor_test.py:41-68 (expiry 0 included: a deterministic caplet), :70-103, floors
  cap = dict(strikes=0.01 * np.ones(4), expiries=np.array([0.0, 0.25, 0.5, 0.75]),
             maturities=np.array([0.25, 0.5, 0.75, 1.0]), daycount_fractions=0.25 * np.ones(4),
             notional=100.0, reference_rate_fn=RATE, num_samples=100_000, time_step=0.1, seed=[42, 42])
  got = tff.models.hjm.cap_floor_price(random_type=rt, dtype=np.float64, **one, **cap)
  want = ohjm.cap_floor_price_mc(random_type=ort, **one, **cap)
  assert got.shape == () and got.dtype == np.float64
  np.testing.assert_allclose(got, want, rtol=1e-10)
  assert abs(got - 0.4072088281493774) < 2.5e-3
  got = tff.models.hjm.cap_floor_price(random_type=rt, dtype=np.float64, is_cap=False, **two, **cap)
  want = ohjm.cap_floor_price_mc(random_type=ort, is_cap=False, **two, **cap)
  np.testing.assert_allclose(got, want, rtol=1e-10)
  from tff_b200.math import piecewise
  pw = piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  opw = omodels.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  got = tff.models.hjm.cap_floor_price(random_type=rt, dtype=np.float64, dim=1, mean_reversion=[0.03],
                                       volatility=lambda t, r: pw([float(t)]), **cap)
  want = ohjm.cap_floor_price_mc(random_type=ort, dim=1, mean_reversion=[0.03],
                                 volatility=lambda t, r: opw(np.asarray([t])), **cap)
  np.testing.assert_allclose(got, want, rtol=1e-10)
  assert abs(got - 0.2394242699989869) < 2.5e-3
  with pytest.raises(ValueError):
    tff.models.hjm.bond_option_price(strikes=strikes, expiries=exp, maturities=mat,
                                     discount_rate_fn=RATE, random_type=rt, **one)
