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
