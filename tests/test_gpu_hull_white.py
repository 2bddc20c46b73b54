"""GPU parity of the Hull-White path (exact OU step, Euler form, discount-curve
paths, fused swaption pricer) against the oracle; tolerances as in
tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import hull_white as ohw
from oracle import models as omodels

pytestmark = pytest.mark.gpu


def _flat(t):
  return 0.01 + 0 * t


def _curve(t):          # analytic, non-flat zero curve
  return 0.01 + 0.002 * t


def _np(t):
  return t.detach().cpu().numpy()


RNGS = [('STATELESS', [4, 2], 0), ('STATELESS_ANTITHETIC', [4, 2], 0),
        ('SOBOL', None, 1000), ('PSEUDO', 7, 0)]


@pytest.mark.parametrize('rng', RNGS, ids=lambda r: r[0])
@pytest.mark.parametrize('curve', [_flat, _curve], ids=['flat', 'sloped'])
def test_hw_exact_paths_match_oracle(rng, curve):
  import tff_b200 as tff
  from tff_b200.math import piecewise
  rt, seed, skip = rng
  dtype = np.float64
  vol = piecewise.PiecewiseConstantFunc([0.1, 0.7], [0.01, 0.02, 0.015], dtype=dtype)
  ovol = omodels.PiecewiseConstantFunc([0.1, 0.7], [0.01, 0.02, 0.015], dtype=dtype)
  model = tff.models.HullWhiteModel1F(0.1, vol, curve, dtype=dtype)
  omodel = ohw.HullWhiteModel1F(0.1, ovol, curve, dtype)
  times = [0.1, 0.5, 1.0, 2.0]
  n = 2000
  got = _np(model.sample_paths(times, num_samples=n,
                               random_type=tff.math.random.RandomType[rt],
                               seed=seed, skip=skip))
  want = omodel.sample_paths(times, n, odraws.RandomType[rt], seed=seed, skip=skip)
  assert got.shape == want.shape == (n, 4, 1)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_hw_single_time_and_times_grid():
  import tff_b200 as tff
  dtype = np.float64
  model = tff.models.HullWhiteModel1F(0.03, 0.02, _flat, dtype=dtype)
  omodel = ohw.HullWhiteModel1F(0.03, 0.02, _flat, dtype)
  rt = tff.math.random.RandomType.STATELESS
  got = _np(model.sample_paths([1.0], num_samples=500, random_type=rt, seed=[1, 2]))
  want = omodel.sample_paths([1.0], 500, odraws.RandomType.STATELESS, seed=[1, 2])
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)
  grid = np.linspace(0, 1, 11)
  got = _np(model.sample_paths([0.32, 0.9], num_samples=500, random_type=rt,
                               seed=[1, 2], times_grid=grid))
  want = omodel.sample_paths([0.32, 0.9], 500, odraws.RandomType.STATELESS,
                             seed=[1, 2], times_grid=grid)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_hw_euler_for_generic_parameters():
  # a generic callable volatility switches the reference to the Euler scheme
  # with initial_state = f(0, 0) (vector_hull_white.py:406-433)
  import tff_b200 as tff
  dtype = np.float64
  a = 0.1

  def vol_fn(t):
    return 0.01 + 0.005 * np.asarray(t)
  model = tff.models.HullWhiteModel1F(a, vol_fn, _curve, dtype=dtype)
  assert model._sample_with_generic
  times = [0.5, 1.0]
  n = 1500
  got = _np(model.sample_paths(times, num_samples=n, time_step=0.05,
                               random_type=tff.math.random.RandomType.SOBOL, skip=3))
  fwd, fwd_grad = omodels.complex_step_forward_rate(_curve)
  # f'(0,t) of the sloped curve is exactly 0.004
  d, v = omodels.hull_white_1f_closures(a, vol_fn, fwd, lambda t: 0.004, dtype)
  want = oeuler.sample(1, d, v, times, time_step=0.05, num_samples=n,
                       initial_state=np.array([fwd(0.0)]),
                       random_type=odraws.RandomType.SOBOL, skip=3, dtype=dtype)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_hw_discount_curve_paths():
  import tff_b200 as tff
  dtype = np.float64
  model = tff.models.HullWhiteModel1F(0.03, 0.02, _curve, dtype=dtype)
  omodel = ohw.HullWhiteModel1F(0.03, 0.02, _curve, dtype)
  times, curve_times = [0.25, 0.5, 1.0], [0.25, 0.5, 0.75, 1.0]
  p, r = model.sample_discount_curve_paths(
      times, curve_times, num_samples=700,
      random_type=tff.math.random.RandomType.STATELESS, seed=[3, 4])
  op, orr = omodel.sample_discount_curve_paths(
      times, curve_times, 700, odraws.RandomType.STATELESS, seed=[3, 4])
  assert tuple(p.shape) == op.shape == (700, 4, 3, 1)
  np.testing.assert_allclose(_np(r), orr, rtol=1e-12, atol=1e-14)
  np.testing.assert_allclose(_np(p), op, rtol=1e-12)


SWAPTION = dict(
    fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
    fixed_leg_daycount_fractions=0.25 * np.ones(4),
    fixed_leg_coupon=0.011 * np.ones(4), mean_reversion=0.03, volatility=0.02)


def _legs():
  return dict(floating_leg_start_times=np.array([1.0, 1.25, 1.5, 1.75]),
              floating_leg_end_times=np.array([1.25, 1.5, 1.75, 2.0]),
              floating_leg_daycount_fractions=0.25 * np.ones(4))


@pytest.mark.parametrize('rng', [('STATELESS', [4, 2]), ('STATELESS_ANTITHETIC', [4, 2]),
                                 ('SOBOL', None)], ids=lambda r: r[0])
def test_swaption_price_matches_oracle(rng):
  import tff_b200 as tff
  rt, seed = rng
  n = 1 << 16
  got = tff.models.hull_white.swaption_price(
      expiries=np.array(1.0), reference_rate_fn=_flat, notional=100.,
      use_analytic_pricing=False, num_samples=n, time_step=0.1,
      random_type=tff.math.random.RandomType[rt], seed=seed, dtype=np.float64,
      **SWAPTION, **_legs())
  want = ohw.swaption_price_mc(
      expiries=np.array(1.0), reference_rate_fn=_flat, notional=100.,
      num_samples=n, time_step=0.1, random_type=odraws.RandomType[rt], seed=seed,
      dtype=np.float64, **SWAPTION)
  assert got.shape == () and got.dtype == np.float64
  np.testing.assert_allclose(got, want, rtol=1e-12)


def test_swaption_reference_kat():
  # models/hull_white/swaption_test.py:85-125: 0.71632434 +- 1e-3 with 500k
  # STATELESS_ANTITHETIC paths, seed [4, 2], time_step 0.1
  import tff_b200 as tff
  price = tff.models.hull_white.swaption_price(
      expiries=np.array(1.0), reference_rate_fn=_flat, notional=100.,
      use_analytic_pricing=False, num_samples=500000, time_step=0.1,
      random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, seed=[4, 2],
      dtype=np.float64, **SWAPTION, **_legs())
  np.testing.assert_allclose(price, 0.71632434, rtol=1e-3, atol=1e-3)


def test_swaption_batch_with_different_expiries():
  # models/hull_white/swaption_test.py:291-318 style batch (payer / receiver,
  # two expiries): mid-path payoff evaluation.  time_step = 0.3 keeps the uniform
  # grid off the expiries: exact duplicates in sim_times shift the reference's
  # own TensorArray slots (a reference quirk the fused pricer does not have).
  import tff_b200 as tff
  expiries = np.array([1.0, 2.0, 1.0])
  pay = np.array([[1.25, 1.5, 1.75, 2.0], [2.25, 2.5, 2.75, 3.0], [1.25, 1.5, 1.75, 2.0]])
  kw = dict(fixed_leg_payment_times=pay, fixed_leg_daycount_fractions=0.25 * np.ones_like(pay),
            fixed_leg_coupon=0.011 * np.ones_like(pay), mean_reversion=0.03, volatility=0.02,
            notional=np.array([100., 50., 100.]), is_payer_swaption=np.array([True, True, False]))
  n = 1 << 15
  got = tff.models.hull_white.swaption_price(
      expiries=expiries, reference_rate_fn=_curve, use_analytic_pricing=False,
      num_samples=n, time_step=0.3, random_type=tff.math.random.RandomType.STATELESS,
      seed=[9, 9], dtype=np.float64, floating_leg_start_times=pay - 0.25,
      floating_leg_end_times=pay, floating_leg_daycount_fractions=0.25 * np.ones_like(pay), **kw)
  want = ohw.swaption_price_mc(
      expiries=expiries, reference_rate_fn=_curve, num_samples=n, time_step=0.3,
      random_type=odraws.RandomType.STATELESS, seed=[9, 9], dtype=np.float64, **kw)
  assert got.shape == (3,)
  np.testing.assert_allclose(got, want, rtol=1e-12)


# ---------------------------------------------------------------- bond options
@pytest.mark.parametrize('rng', [('STATELESS', [4, 2]), ('STATELESS_ANTITHETIC', [4, 2]),
                                 ('SOBOL', None)], ids=lambda r: r[0])
def test_bond_option_matches_oracle(rng):
  import tff_b200 as tff
  from tff_b200.math import piecewise
  rt, seed = rng
  n = 1 << 15
  strikes = np.array([[0.95, 0.97], [0.93, 0.99]])
  expiries = np.array([[1.0, 1.0], [0.55, 2.0]])
  maturities = np.array([[5.0, 3.0], [4.0, 2.5]])
  is_call = np.array([[True, False], [False, True]])
  vol = piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  ovol = omodels.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  got = tff.models.hull_white.bond_option_price(
      strikes=strikes, expiries=expiries, maturities=maturities,
      discount_rate_fn=_curve, mean_reversion=0.03, volatility=vol,
      is_call_options=is_call, use_analytic_pricing=False, num_samples=n,
      time_step=0.1, random_type=tff.math.random.RandomType[rt], seed=seed,
      dtype=np.float64)
  want = ohw.bond_option_price_mc(
      strikes=strikes, expiries=expiries, maturities=maturities,
      discount_rate_fn=_curve, mean_reversion=0.03, volatility=ovol,
      is_call_options=is_call, num_samples=n, time_step=0.1,
      random_type=odraws.RandomType[rt], seed=seed)
  assert got.shape == (2, 2) and got.dtype == np.float64
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-16)


def test_bond_option_reference_kat():
  # models/hull_white/zero_coupon_bond_option_test.py:49-73 (0.02817777 +- 1e-4,
  # 500k antithetic paths) and :119-146 (time-dependent vol, 0.02237839); the
  # dt = 0.1 discount-factor quadrature alone biases the price by +5.5e-5
  import tff_b200 as tff
  from tff_b200.math import piecewise
  expiries, maturities = np.array(1.0), np.array(5.0)
  strikes = np.exp(-0.01 * maturities) / np.exp(-0.01 * expiries)
  kw = dict(strikes=strikes, expiries=expiries, maturities=maturities,
            mean_reversion=0.03, discount_rate_fn=_flat, use_analytic_pricing=False,
            num_samples=500000, time_step=0.1,
            random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, seed=[1, 7],
            dtype=np.float64)
  price = tff.models.hull_white.bond_option_price(volatility=0.02, **kw)
  assert price.shape == ()
  np.testing.assert_allclose(price, 0.02817777, rtol=1e-4, atol=1e-4)
  vol = piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  price = tff.models.hull_white.bond_option_price(volatility=vol, **kw)
  np.testing.assert_allclose(price, 0.02237839, rtol=1e-4, atol=1e-4)


def test_bond_option_errors():
  import tff_b200 as tff
  kw = dict(strikes=np.array(0.9), expiries=np.array(1.0), maturities=np.array(2.0),
            discount_rate_fn=_flat, mean_reversion=0.03, volatility=0.02,
            dtype=np.float64)
  with pytest.raises(ValueError, match='time_step'):
    tff.models.hull_white.bond_option_price(use_analytic_pricing=False, **kw)
  # the default (analytic) valuation is a host-side closed form: MC at 1e-3 of it
  analytic = tff.models.hull_white.bond_option_price(use_analytic_pricing=True, **kw)
  mc = tff.models.hull_white.bond_option_price(
      use_analytic_pricing=False, num_samples=400000, time_step=0.1, seed=[1, 2],
      random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, **kw)
  np.testing.assert_allclose(mc, analytic, rtol=0, atol=2e-4)


# ----------------------------------------------------------------- caps/floors
CAP = dict(expiries=np.array([0.0, 0.25, 0.5, 0.75]),
           maturities=np.array([0.25, 0.5, 0.75, 1.0]),
           strikes=0.01 * np.ones(4), daycount_fractions=0.25 * np.ones(4))


def test_cap_price_reference_kat_and_oracle():
  # models/hull_white/cap_floor_test.py:57-83: 0.4072088281493774 +- 1e-3 with
  # 50k STATELESS_ANTITHETIC paths, seed [42, 42] (first caplet expires at t=0)
  import tff_b200 as tff
  kw = dict(notional=100.0, mean_reversion=0.03, volatility=0.02,
            reference_rate_fn=_flat, num_samples=50_000, time_step=0.1,
            seed=[42, 42], dtype=np.float64, **CAP)
  price = tff.models.hull_white.cap_floor_price(
      use_analytic_pricing=False,
      random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, **kw)
  assert price.shape == () and price.dtype == np.float64
  np.testing.assert_allclose(price, 0.4072088281493774, rtol=1e-3, atol=1e-3)
  want = ohw.cap_floor_price_mc(
      random_type=odraws.RandomType.STATELESS_ANTITHETIC, **kw)
  np.testing.assert_allclose(price, want, rtol=1e-12)


def test_cap_floor_batch_matches_oracle():
  # cap_floor_test.py:228-263 style 2-d batch: caps and floors, two strikes,
  # piecewise-constant volatility
  import tff_b200 as tff
  from tff_b200.math import piecewise
  expiries = np.broadcast_to(CAP['expiries'], (2, 2, 4))
  maturities = np.broadcast_to(CAP['maturities'], (2, 2, 4))
  strikes = np.array([[0.01, 0.02], [0.01, 0.02]])[..., None] * np.ones(4)
  dcf = 0.25 * np.ones((2, 2, 4))
  is_cap = np.array([[True, True], [False, False]])[..., None]
  vol = piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  ovol = omodels.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  kw = dict(strikes=strikes, expiries=expiries, maturities=maturities,
            daycount_fractions=dcf, notional=100.0, mean_reversion=0.03,
            reference_rate_fn=_curve, is_cap=is_cap, num_samples=1 << 15,
            time_step=0.1, seed=[3, 5], dtype=np.float64)
  got = tff.models.hull_white.cap_floor_price(
      volatility=vol, use_analytic_pricing=False,
      random_type=tff.math.random.RandomType.STATELESS, **kw)
  want = ohw.cap_floor_price_mc(
      volatility=ovol, random_type=odraws.RandomType.STATELESS, **kw)
  assert got.shape == (2, 2)
  np.testing.assert_allclose(got, want, rtol=1e-12)


def test_cap_reference_kat_time_dependent_vol():
  # cap_floor_test.py:142-170: 0.2394242699989869 +- 1e-2
  import tff_b200 as tff
  from tff_b200.math import piecewise
  vol = piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  price = tff.models.hull_white.cap_floor_price(
      notional=100.0, mean_reversion=0.03, volatility=vol, reference_rate_fn=_flat,
      use_analytic_pricing=False, num_samples=100_000, time_step=0.1,
      random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, seed=[42, 42],
      dtype=np.float64, **CAP)
  np.testing.assert_allclose(price, 0.2394242699989869, rtol=1e-2, atol=1e-2)


# ------------------------------------------------------ correlated factors ----
def _flat2(t):
  import torch
  if isinstance(t, torch.Tensor):
    return 0.01 + 0 * t.unsqueeze(-1) * torch.ones(2, dtype=t.dtype)
  return 0.01 + 0 * np.asarray(t)[..., None] * np.ones(2)


def _curve2(t):
  import torch
  if isinstance(t, torch.Tensor):
    return 0.01 + t.unsqueeze(-1) * torch.tensor([0.002, 0.001], dtype=t.dtype)
  return 0.01 + np.asarray(t)[..., None] * np.array([0.002, 0.001])


def _vector_models(curve, corr_kind, dtype=np.float64):
  import tff_b200 as tff
  from tff_b200.math import piecewise
  mr = [0.1, 0.05]
  jumps = [[0.1, 0.2, 0.5], [0.1, 2.0, 3.0]]
  vals = [[0.01, 0.012, 0.011, 0.013], [0.02, 0.018, 0.02, 0.02]]
  vol = piecewise.PiecewiseConstantFunc(jumps, vals, dtype=dtype)
  ovol = [omodels.PiecewiseConstantFunc(jumps[i], vals[i], dtype=dtype) for i in range(2)]
  if corr_kind == 'none':
    corr = ocorr = None
  elif corr_kind == 'const':
    corr = ocorr = [[1., 0.5], [0.5, 1.]]
  else:
    cv = [[[1., 0.5], [0.5, 1.]], [[1., 0.6], [0.6, 1.]], [[1., 0.9], [0.9, 1.]]]
    corr = piecewise.PiecewiseConstantFunc([0.5, 2.0], cv, dtype=dtype)
    ocorr = omodels.PiecewiseConstantFunc([0.5, 2.0], cv, dtype=dtype)
  model = tff.models.hull_white.VectorHullWhiteModel(
      2, mr, vol, curve, corr_matrix=corr, dtype=dtype)
  omodel = ohw.VectorHullWhiteModel(2, mr, ovol, curve, ocorr, dtype)
  return model, omodel


@pytest.mark.parametrize('rng', RNGS, ids=lambda r: r[0])
@pytest.mark.parametrize('corr_kind', ['none', 'const', 'piecewise'])
def test_vector_hw_paths_match_oracle(rng, corr_kind):
  import tff_b200 as tff
  rt, seed, skip = rng
  model, omodel = _vector_models(_curve2, corr_kind)
  times = [0.1, 0.5, 1.0, 2.5]
  n = 2000
  got = _np(model.sample_paths(times, num_samples=n,
                               random_type=tff.math.random.RandomType[rt],
                               seed=seed, skip=skip))
  want = omodel.sample_paths(times, n, odraws.RandomType[rt], seed=seed, skip=skip)
  assert got.shape == want.shape == (n, 4, 2)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize('corr_kind', ['none', 'piecewise'])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_vector_hw_discount_curve_paths_match_oracle(corr_kind, dtype):
  # vector_hull_white.py:451-592, 783-814: [N, m, k, dim] bond prices from the kernel
  import tff_b200 as tff
  model, omodel = _vector_models(_curve2, corr_kind, dtype)
  times, curve_times = [0.1, 0.5, 1.0, 2.5], [0.25, 0.5, 1.0]
  n = 1531                      # not a multiple of the 32-path tile
  p, r = model.sample_discount_curve_paths(
      times, curve_times, num_samples=n, random_type=tff.math.random.RandomType.STATELESS,
      seed=[5, 6])
  op, orr = ohw.vector_sample_discount_curve_paths(
      omodel, times, curve_times, n, odraws.RandomType.STATELESS, seed=[5, 6])
  assert tuple(p.shape) == op.shape == (n, 3, 4, 2) and tuple(r.shape) == (n, 4, 2)
  tol = 1e-12 if dtype == np.float64 else 1e-5
  np.testing.assert_allclose(_np(r), orr, rtol=tol, atol=tol * 1e-2)
  np.testing.assert_allclose(_np(p), op, rtol=tol)


def test_hw_discount_curve_paths_many_times_and_strided_rates():
  # 1-factor: 40 simulation times x 5 maturities (several 32-column tiles)
  import tff_b200 as tff
  dtype = np.float64
  model = tff.models.HullWhiteModel1F(0.03, 0.02, _curve, dtype=dtype)
  omodel = ohw.HullWhiteModel1F(0.03, 0.02, _curve, dtype)
  times = np.linspace(0.05, 2.0, 40)
  curve_times = [0.25, 0.5, 1.0, 2.0, 5.0]
  p, r = model.sample_discount_curve_paths(
      times, curve_times, num_samples=333, random_type=tff.math.random.RandomType.SOBOL, skip=7)
  op, orr = omodel.sample_discount_curve_paths(times, curve_times, 333, odraws.RandomType.SOBOL,
                                               skip=7)
  assert tuple(p.shape) == op.shape == (333, 5, 40, 1)
  np.testing.assert_allclose(_np(p), op, rtol=1e-12)


def test_vector_hw_reference_moments_kat():
  # hull_white_test.py:223-270: mean / variance (1e-4) and correlation (1e-2)
  # at t = 1 with 50k STATELESS_ANTITHETIC paths, seed [1, 2]
  import tff_b200 as tff
  from tff_b200.math import piecewise
  a, sigma = np.array([0.1, 0.05]), np.array([0.01, 0.02])
  vol = piecewise.PiecewiseConstantFunc(
      [[0.1, 0.2, 0.5], [0.1, 2.0, 3.0]], [4 * [sigma[0]], 4 * [sigma[1]]], dtype=np.float64)
  model = tff.models.hull_white.VectorHullWhiteModel(
      2, a, vol, _flat2, corr_matrix=[[1., 0.5], [0.5, 1.]], dtype=np.float64)
  paths = _np(model.sample_paths(
      [0.1, 0.5, 1.0], num_samples=50000,
      random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, seed=[1, 2]))
  assert paths.shape == (50000, 3, 2) and paths.dtype == np.float64
  x = paths[:, -1, :]
  np.testing.assert_allclose(x.mean(0), 0.01 + sigma**2 / 2 / a**2 * (1 - np.exp(-a))**2,
                             rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(x.var(0), sigma**2 / 2 / a * (1 - np.exp(-2 * a)),
                             rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(np.corrcoef(x[:, 0], x[:, 1])[0, 1], 0.5, atol=1e-2)


def test_vector_hw_supplied_draws_and_grid():
  # hull_white_test.py:283-342 ("SupplyDrawsSupplyGrid"): user draws [N, steps, 2]
  # on a user grid
  import torch
  import tff_b200 as tff
  model, omodel = _vector_models(_flat2, 'piecewise')
  grid = [0.0, 0.1, 0.2, 0.5, 1.0]
  n = 1000
  half = np.random.RandomState(3).standard_normal((n // 2, 4, 2))
  draws = np.concatenate([half, -half], 0)
  got = _np(model.sample_paths([0.1, 0.5, 1.0], normal_draws=torch.as_tensor(draws).cuda(),
                               times_grid=grid))
  want = omodel.sample_paths([0.1, 0.5, 1.0], n, normal_draws=draws, times_grid=grid)
  assert got.shape == (n, 3, 2)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_vector_hw_generic_corr_uses_euler():
  # hull_white_test.py:344-383: a generic callable correlation -> Euler scheme
  # on the model closures; parity with the oracle Euler sampler on the same
  # closures
  import torch
  import tff_b200 as tff
  a, sigma = np.array([0.1, 0.05]), np.array([0.01, 0.02])

  def corr(t):
    one = torch.ones((), dtype=torch.float64) if isinstance(t, torch.Tensor) else 1.0
    rho = 0.5 * one
    if isinstance(t, torch.Tensor):
      return torch.stack([torch.stack([one, rho]), torch.stack([rho, one])])
    return np.array([[1.0, 0.5], [0.5, 1.0]])
  model = tff.models.hull_white.VectorHullWhiteModel(
      2, a, sigma, _flat2, corr_matrix=corr, dtype=np.float64)
  with pytest.raises(ValueError, match='time_step'):
    model.sample_paths([0.5, 1.0], num_samples=10)
  n = 4000
  got = _np(model.sample_paths([0.5, 1.0], num_samples=n, time_step=0.1,
                               random_type=tff.math.random.RandomType.STATELESS,
                               seed=[7, 1]))
  root = np.linalg.cholesky(np.array([[1.0, 0.5], [0.5, 1.0]]))

  def drift_fn(t, x):
    return a * 0.01 + sigma**2 / 2 / a * (1 - np.exp(-2 * a * t)) - a * x

  def vol_fn(t, x):
    return np.broadcast_to(sigma[:, None] * root, x.shape[:-1] + (2, 2))
  want = oeuler.sample(2, drift_fn, vol_fn, [0.5, 1.0], time_step=0.1, num_samples=n,
                       initial_state=np.array([0.01, 0.01]),
                       random_type=odraws.RandomType.STATELESS, seed=[7, 1],
                       dtype=np.float64)
  assert got.shape == (n, 2, 2)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_vector_hw_dim1_and_errors():
  import tff_b200 as tff
  m1 = tff.models.hull_white.VectorHullWhiteModel(1, [0.1], [0.01], _flat, dtype=np.float64)
  m0 = tff.models.HullWhiteModel1F(0.1, 0.01, _flat, dtype=np.float64)
  kw = dict(num_samples=500, random_type=tff.math.random.RandomType.STATELESS, seed=[1, 2])
  np.testing.assert_array_equal(_np(m1.sample_paths([0.5, 1.0], **kw)),
                                _np(m0.sample_paths([0.5, 1.0], **kw)))
  with pytest.raises(ValueError, match='should be the same as `dims`'):
    tff.models.hull_white.VectorHullWhiteModel(2, [0.1], [0.01, 0.02], _flat2, dtype=np.float64)
  m2 = tff.models.hull_white.VectorHullWhiteModel(2, [0.1, 0.2], [0.01, 0.02], _flat2,
                                                  dtype=np.float64)
  with pytest.raises(ValueError, match='rank 1'):
    m2.sample_paths([[0.5, 1.0]], num_samples=10)


# ------------------------------------------------------- Bermudan swaptions ----
def _bermudan_legs(exercise):
  start = np.array([np.clip(np.arange(8) * 0.5 + e, None, 5.0) for e in exercise])
  end = np.clip(start + 0.5, 0.0, 5.0)
  return dict(exercise_times=np.array(exercise), fixed_leg_payment_times=end,
              fixed_leg_daycount_fractions=end - start,
              fixed_leg_coupon=0.011 * np.ones_like(end))


EX1 = [1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5]
EX2 = [2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 5.0]


def _product_legs(legs):
  return dict(floating_leg_start_times=legs['fixed_leg_payment_times'] -
              legs['fixed_leg_daycount_fractions'],
              floating_leg_end_times=legs['fixed_leg_payment_times'],
              floating_leg_daycount_fractions=legs['fixed_leg_daycount_fractions'], **legs)


def test_bermudan_swaption_reference_kat_and_oracle():
  # bermudan_swaption_test.py:86-131: 1.8892 +- 1e-2 (10k paths, seed [0, 0])
  import tff_b200 as tff
  legs = _bermudan_legs(EX1)
  kw = dict(reference_rate_fn=_flat, notional=100., mean_reversion=0.03, volatility=0.01,
            num_samples=10000, time_step=0.1, seed=[0, 0], dtype=np.float64)
  got = tff.models.hull_white.bermudan_swaption_price(
      random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, **_product_legs(legs), **kw)
  assert got.shape == () and got.dtype == np.float64
  np.testing.assert_allclose(got, 1.8892, rtol=1e-2, atol=1e-2)
  want = ohw.bermudan_swaption_price_mc(
      random_type=odraws.RandomType.STATELESS_ANTITHETIC, **legs, **kw)
  np.testing.assert_allclose(got, want, rtol=1e-9)


@pytest.mark.parametrize('rng', [('STATELESS_ANTITHETIC', [0, 0]), ('SOBOL', None)],
                         ids=lambda r: r[0])
def test_bermudan_swaption_batch_matches_oracle(rng):
  # bermudan_swaption_test.py:133-183: [5nc1, 5nc2] = [1.8892, 1.6633] +- 5e-3;
  # piecewise-constant volatility and a sloped curve for the parity leg
  import tff_b200 as tff
  from tff_b200.math import piecewise
  rt, seed = rng
  l1, l2 = _bermudan_legs(EX1), _bermudan_legs(EX2)
  legs = {k: np.stack([l1[k], l2[k]]) for k in l1}
  kw = dict(notional=100., mean_reversion=0.03, num_samples=50000, time_step=0.1,
            seed=seed, dtype=np.float64)
  got = tff.models.hull_white.bermudan_swaption_price(
      reference_rate_fn=_flat, volatility=0.01,
      random_type=tff.math.random.RandomType[rt], **_product_legs(legs), **kw)
  assert got.shape == (2,)
  np.testing.assert_allclose(got, [1.8892, 1.6633], rtol=5e-3, atol=5e-3)
  vol = piecewise.PiecewiseConstantFunc([2.0], [0.01, 0.012], dtype=np.float64)
  ovol = omodels.PiecewiseConstantFunc([2.0], [0.01, 0.012], dtype=np.float64)
  kw['num_samples'] = 20000
  got = tff.models.hull_white.bermudan_swaption_price(
      reference_rate_fn=_curve, volatility=vol,
      random_type=tff.math.random.RandomType[rt], **_product_legs(legs), **kw)
  want = ohw.bermudan_swaption_price_mc(
      reference_rate_fn=_curve, volatility=ovol, random_type=odraws.RandomType[rt],
      **legs, **kw)
  np.testing.assert_allclose(got, want, rtol=1e-9)


def test_bermudan_swaption_errors():
  import tff_b200 as tff
  legs = _product_legs(_bermudan_legs(EX1))
  kw = dict(reference_rate_fn=_flat, mean_reversion=0.03, volatility=0.01, dtype=np.float64)
  with pytest.raises(ValueError, match='time_step'):
    tff.models.hull_white.bermudan_swaption_price(**legs, **kw)
  with pytest.raises(NotImplementedError):
    tff.models.hull_white.bermudan_swaption_price(use_finite_difference=True, **legs, **kw)
