"""HJM models and their Monte-Carlo pricers: the HOST flow of the public calls on the CPU.

The same calls as the GPU tests of `tests/test_hjm.py`, with the device plan replaced by
`tests/cpu_plan.CpuPlan` (`HjmModel::step` and the multi-factor swaption payoff restated in numpy;
installed by pytest's `monkeypatch` inside the test): everything the mirror does on the host -- the
two grids, the deterministic y tables, the per-step coefficient table, the reference's draw layout
(the quasi-Gaussian state consumes F + F^2 normals per step), payoff descriptors, batches, expiry-0
caplets -- against the oracle (`oracle/hjm.py`, pinned by the reference's values in `tests/test_hjm.py`).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cpu_plan  # pylint: disable=g-import-not-at-top

from oracle import draws as odraws
from oracle import hjm as ohjm
from oracle import models as omodels
import tff_b200 as tff

RATE = lambda t: 0.01 + 0 * t
SWAPTION = dict(expiries=np.array([1.0]), fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
                fixed_leg_daycount_fractions=0.25 * np.ones(4), fixed_leg_coupon=0.011 * np.ones(4),
                reference_rate_fn=RATE, notional=100., seed=[1, 2], dtype=np.float64)
RT, ORT = tff.math.random.RandomType.STATELESS_ANTITHETIC, odraws.RandomType.STATELESS_ANTITHETIC


@pytest.fixture
def cpu_engine(monkeypatch):
  cpu_plan.install(monkeypatch)


def _np(t):
  return t.detach().cpu().numpy()


@pytest.mark.parametrize('factors,grid', [(1, dict(time_step=0.1)), (1, dict(num_time_steps=11)), (2, dict(time_step=0.05))])
@pytest.mark.parametrize('rt_name', ['STATELESS_ANTITHETIC', 'SOBOL'])
def test_quasi_gaussian_sample_paths(cpu_engine, factors, grid, rt_name):
  mr, vol = [0.03, 0.06][:factors], [0.02, 0.01][:factors]
  corr = None if factors == 1 else [[1.0, 0.4], [0.4, 1.0]]
  times = np.array([0.3, 1.0, 1.7, 1.7])           # a repeated time
  model = tff.models.hjm.QuasiGaussianHJM(factors, mr, vol, RATE, corr_matrix=corr, dtype=np.float64)
  want = ohjm.QuasiGaussianHJM(factors, mr, vol, RATE, corr_matrix=corr)
  kw = dict(seed=[4, 2], skip=3, **grid)
  rate, df, x, y = model.sample_paths(times, 200, random_type=getattr(tff.math.random.RandomType, rt_name), **kw)
  wr, wdf, wx, wy = want.sample_paths(times, 200, random_type=getattr(odraws.RandomType, rt_name), **kw)
  np.testing.assert_allclose(_np(x), wx, rtol=1e-11, atol=1e-15)
  np.testing.assert_allclose(_np(rate), wr, rtol=1e-11, atol=1e-15)
  np.testing.assert_allclose(_np(df), wdf, rtol=1e-12)
  np.testing.assert_allclose(_np(y), wy, rtol=1e-12, atol=1e-20)


@pytest.mark.parametrize('factors,mr,vol,corr,grid', [
    (1, [0.03], [0.01], None, dict(num_time_steps=21)),
    (2, [0.03, 0.1], [0.005, 0.012], [[1.0, 0.5], [0.5, 1.0]], dict(time_step=0.1)),
    (3, [0.03, 0.1, 0.2], [0.005, 0.012, 0.007], None, dict(time_step=0.1))])
def test_gaussian_hjm_sample_paths(cpu_engine, factors, mr, vol, corr, grid):
  times = np.array([0.1, 0.5, 1.0, 2.0])
  model = tff.models.hjm.GaussianHJM(factors, mr, vol, RATE, corr_matrix=corr, dtype=np.float64)
  want = ohjm.GaussianHJM(factors, mr, vol, RATE, corr_matrix=corr)
  rate, df, x, y = model.sample_paths(times, 200, random_type=RT, seed=[1, 2], **grid)
  wr, wdf, wx, wy = want.sample_paths(times, 200, random_type=ORT, seed=[1, 2], **grid)
  np.testing.assert_allclose(_np(x), wx, rtol=1e-11, atol=1e-15)
  np.testing.assert_allclose(_np(rate), wr, rtol=1e-11, atol=1e-15)
  np.testing.assert_allclose(_np(df), wdf, rtol=1e-12)
  np.testing.assert_allclose(_np(y), wy, rtol=1e-12, atol=1e-20)


ONE = dict(num_hjm_factors=1, mean_reversion=[0.03], volatility=[0.02])
TWO = dict(num_hjm_factors=2, mean_reversion=[0.03, 0.06], volatility=[0.02, 0.01])


@pytest.mark.parametrize('model_kw,grid', [
    (ONE, dict(time_step=0.1)), (ONE, dict(num_time_steps=11)), (dict(ONE, is_payer_swaption=False), dict(time_step=0.1)),
    (TWO, dict(time_step=0.1)), (dict(TWO, corr_matrix=[[1.0, 0.5], [0.5, 1.0]]), dict(time_step=0.1))])
def test_hjm_swaption_price(cpu_engine, model_kw, grid):
  # swaption_pricing_test.py:46-165, 321-356
  got = tff.models.hjm.swaption_price(num_samples=2000, random_type=RT, **SWAPTION, **model_kw, **grid)
  want = ohjm.swaption_price_mc(num_samples=2000, random_type=ORT, **SWAPTION, **model_kw, **grid)
  assert got.shape == (1,) and got.dtype == np.float64
  np.testing.assert_allclose(got, want, rtol=1e-10)


def test_hjm_swaption_batch_and_callable_volatility(cpu_engine):
  kw = dict(expiries=np.array([1.0, 2.0, 1.0]),
            fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0], [2.25, 2.5, 2.75, 3.0], [1.25, 1.5, 1.75, 2.0]]),
            fixed_leg_daycount_fractions=0.25 * np.ones((3, 4)), fixed_leg_coupon=0.011 * np.ones((3, 4)),
            reference_rate_fn=RATE, notional=np.array([100., 50., 100.]),
            is_payer_swaption=np.array([True, True, False]), seed=[1, 2], dtype=np.float64, num_samples=1000,
            time_step=0.1)
  np.testing.assert_allclose(tff.models.hjm.swaption_price(random_type=RT, **ONE, **kw),
                             ohjm.swaption_price_mc(random_type=ORT, **ONE, **kw), rtol=1e-10)
  pw = tff.math.piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  opw = omodels.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  got = tff.models.hjm.swaption_price(num_hjm_factors=1, mean_reversion=[0.03], volatility=lambda t, r: pw([float(t)]),
                                      num_samples=2000, random_type=RT, time_step=0.1, **SWAPTION)
  want = ohjm.swaption_price_mc(num_hjm_factors=1, mean_reversion=[0.03], volatility=lambda t, r: opw(np.asarray([t])),
                                num_samples=2000, random_type=ORT, time_step=0.1, **SWAPTION)
  np.testing.assert_allclose(got, want, rtol=1e-10)


def test_hjm_bond_option_and_cap_floor(cpu_engine):
  one = dict(dim=1, mean_reversion=[0.03], volatility=[0.02])
  two = dict(dim=2, mean_reversion=[0.03, 0.06], volatility=[0.02, 0.01])
  exp, mat = np.array([1.0]), np.array([5.0])
  strikes = np.exp(-0.01 * mat) / np.exp(-0.01 * exp)
  for model_kw in (one, two, dict(two, corr_matrix=[[1.0, 0.5], [0.5, 1.0]])):
    kw = dict(strikes=strikes, expiries=exp, maturities=mat, discount_rate_fn=RATE, time_step=0.1, seed=[1, 2],
              num_samples=2000, **model_kw)
    got = tff.models.hjm.bond_option_price(random_type=RT, dtype=np.float64, **kw)
    assert got.shape == (1,) and got.dtype == np.float64
    np.testing.assert_allclose(got, ohjm.bond_option_price_mc(random_type=ORT, **kw), rtol=1e-10)
  kw = dict(strikes=np.array([[0.96, 0.97], [0.99, 0.95]]), expiries=np.array([[1.0, 0.55], [0.25, 1.0]]),
            maturities=np.array([[5.0, 2.0], [0.5, 3.0]]), discount_rate_fn=RATE, time_step=0.1,
            is_call_options=np.array([[True, False], [True, False]]), seed=[4, 2], num_samples=2000)
  got, stderr, bad = tff.models.hjm.bond_option_price(random_type=RT, dtype=np.float64, return_stats=True, **one, **kw)
  assert got.shape == (2, 2) and np.all(bad == 0) and np.all(stderr > 0)
  np.testing.assert_allclose(got, ohjm.bond_option_price_mc(random_type=ORT, **one, **kw), rtol=1e-10)
  cap = dict(strikes=0.01 * np.ones(4), expiries=np.array([0.0, 0.25, 0.5, 0.75]),
             maturities=np.array([0.25, 0.5, 0.75, 1.0]), daycount_fractions=0.25 * np.ones(4), notional=100.0,
             reference_rate_fn=RATE, num_samples=2000, time_step=0.1, seed=[42, 42])
  got = tff.models.hjm.cap_floor_price(random_type=RT, dtype=np.float64, **one, **cap)
  assert got.shape == () and got.dtype == np.float64
  np.testing.assert_allclose(got, ohjm.cap_floor_price_mc(random_type=ORT, **one, **cap), rtol=1e-10)
  got = tff.models.hjm.cap_floor_price(random_type=RT, dtype=np.float64, is_cap=False, **two, **cap)
  np.testing.assert_allclose(got, ohjm.cap_floor_price_mc(random_type=ORT, is_cap=False, **two, **cap), rtol=1e-10)
