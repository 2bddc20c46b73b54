"""The reference's statistical tests of the model samplers, run on the ORACLE (CPU).

`geometric_brownian_motion_test.py` and `heston_model_test.py` pin the exact
log-normal samplers and the QE scheme by moments only; these are those checks
on `oracle/models.py` (`gbm_exact_sample_paths`, `mvgbm_exact_sample_paths`) and
`oracle/heston_qe.py`, with the reference's processes, sample counts,
generators, seeds and tolerances.  The GPU kernels are compared with the same
oracle functions to 1e-12 in `tests/test_gpu_parity.py`.
"""
import numpy as np
import pytest
from scipy import integrate

from oracle import draws as odraws
from oracle import heston_qe as oqe
from oracle import models as omodels
from oracle import philox as ophilox

NUM_SAMPLES = 100000     # geometric_brownian_motion_test.py:28
NUM_STDERRS = 3.0        # :29


def _log_moments(samples, num_samples):
  # geometric_brownian_motion_test_utils.py:133-166
  log_s = np.log(samples)
  mean = log_s.mean(axis=-3, keepdims=True)
  var = ((log_s - mean)**2).mean(axis=-3, keepdims=True)
  mean, var = mean[..., 0, :, 0], var[..., 0, :, 0]
  return mean, var, np.sqrt(var / num_samples), var * np.sqrt(2.0 / (num_samples - 1.0))


def _gbm(mu, sigma, times, dtype):
  # geometric_brownian_motion_test_utils.py:117-131: STATELESS, seed [1234, 5]
  return omodels.gbm_exact_sample_paths(mu, sigma, times, initial_state=[2.0], num_samples=NUM_SAMPLES,
                                        random_type=odraws.RandomType.STATELESS, seed=[1234, 5], dtype=dtype)


def _within(actual, expected, tol):
  # geometric_brownian_motion_test_utils.py:47-91: one tolerance per element
  assert actual.shape == expected.shape == tol.shape
  assert np.all(np.abs(actual - expected) < tol), (actual, expected, tol)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_univariate_sample_mean_and_variance_constant_parameters(dtype):
  # geometric_brownian_motion_test.py:248-269
  mu = sigma = 0.05
  times = np.array([0.1, 0.5, 1.0], dtype=dtype)
  mean, var, se_mean, se_var = _log_moments(_gbm(mu, sigma, times, dtype), NUM_SAMPLES)
  _within(mean, (mu - sigma**2 / 2) * times + np.log(dtype(2.0)), se_mean * NUM_STDERRS)
  _within(var, sigma**2 * times, se_var * NUM_STDERRS)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_univariate_sample_mean_and_variance_time_varying_drift(dtype):
  # geometric_brownian_motion_test.py:428-505
  min_tol = 1e-8 if dtype == np.float64 else 5e-3
  times = np.array([0.0, 1.0, 5.0, 7.0, 10.0], dtype=dtype)
  mu = omodels.PiecewiseConstantFunc(np.array([0.0, 5.0, 10.0], dtype), np.array([0.0, 0.0, 0.05, 0.05], dtype),
                                     dtype=dtype)
  mean, var, se_mean, se_var = _log_moments(_gbm(mu, 0.0, times, dtype), NUM_SAMPLES)
  expected = np.array([0.0, 0.0, 0.0, 2.0 * 0.05, 5.0 * 0.05], dtype) + np.log(dtype(2.0))
  _within(mean, expected, np.maximum(se_mean * NUM_STDERRS, min_tol))
  _within(var, np.zeros(5, dtype), np.maximum(se_var * NUM_STDERRS, min_tol))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_univariate_sample_mean_and_variance_time_varying_vol(dtype):
  # geometric_brownian_motion_test.py:581-621 (and the second half of :428-505)
  min_tol = 1e-8 if dtype == np.float64 else 5e-3
  mu = 0.05
  sigma = omodels.PiecewiseConstantFunc(np.array([0.0, 5.0, 10.0], dtype), np.array([0.0, 0.2, 0.4, 0.6], dtype),
                                        dtype=dtype)
  times = np.array([0.0, 1.0, 5.0, 7.0, 10.0], dtype=dtype)
  mean, var, se_mean, se_var = _log_moments(_gbm(mu, sigma, times, dtype), NUM_SAMPLES)
  expected_var = np.array([0.0, 0.2**2, 5 * 0.2**2, 5 * 0.2**2 + 2 * 0.4**2, 5 * 0.2**2 + 5 * 0.4**2], dtype)
  expected_mean = (times * mu - 0.5 * expected_var + np.log(2.0)).astype(dtype)
  _within(mean, expected_mean, np.maximum(se_mean * NUM_STDERRS, min_tol))
  _within(var, expected_var, np.maximum(se_var * NUM_STDERRS, min_tol))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_multivariate_sample_mean_and_variance(dtype):
  # geometric_brownian_motion_test.py:824-860: SOBOL, 10000 paths
  means, vols = np.array([0.05, 0.05]), np.array([0.1, 0.2])
  corr = [[1, 0.1], [0.1, 1]]
  times = np.array([0.1, 0.5, 1.0])
  x0 = [1.0, 2.0]
  samples = omodels.mvgbm_exact_sample_paths(means, vols, corr, times, initial_state=x0, num_samples=10000,
                                             random_type=odraws.RandomType.SOBOL, dtype=dtype)
  assert samples.shape == (10000, 3, 2) and samples.dtype == dtype
  log_s = np.log(samples.astype(np.float64))
  mean = log_s.mean(axis=0)
  var = ((log_s - mean)**2).mean(axis=0)
  np.testing.assert_allclose(mean, (means - vols**2 / 2) * times[:, None] + np.log(x0), atol=1e-3, rtol=1e-3)
  np.testing.assert_allclose(var, vols**2 * times[:, None], atol=1e-3, rtol=1e-3)
  for i in range(3):
    np.testing.assert_allclose(np.corrcoef(samples[:, i, :], rowvar=False), corr, atol=1e-2, rtol=1e-2)


# ------------------------------------------------------------- Heston QE ----
def test_heston_volatility_stays_at_theta():
  # heston_model_test.py:32-49 (seed=None there: any stream must pass)
  theta = 0.05
  times = np.linspace(0.0, 1.0, 365)
  paths = oqe.sample_paths(1.0, theta, 0.00001, -0.0, times, np.array([np.log(100), 0.045]), num_samples=2,
                           time_step=0.01, random_type=odraws.RandomType.STATELESS, seed=[3, 4])
  assert np.max(np.abs(paths[:, 50:, 1] - theta)) < 0.5e-2


def test_heston_state_behaves_like_gbm():
  # heston_model_test.py:51-77
  times = [0.0, 0.5, 1.0]
  paths = oqe.sample_paths(1.0, 1.0, 0.00001, -0.0, times, np.array([np.log(100), 1.0]), num_samples=1000,
                           time_step=0.001, random_type=odraws.RandomType.STATELESS, seed=[3, 4])
  state = paths[..., 0]
  np.testing.assert_allclose(state[:, 0], np.log(100), 1e-8)
  for i in (1, 2):
    np.testing.assert_allclose(np.mean(np.exp(state[:, i])), 100, 1.0)
    np.testing.assert_allclose(np.std(np.exp(state[:, 1])), 100 * np.sqrt(np.exp(times[i]) - 1), 2.0)


def test_heston_expected_total_variance_mc():
  # heston_model_test.py:97-135; the closed form is `expected_total_variance` (:79-95)
  kappa, theta, v0, future_time = 10.0, 0.04, 0.1, 1.0
  times = np.linspace(0, future_time, 252)
  paths = oqe.sample_paths(kappa, theta, 1.0, -0.5, times, np.array([1.0, v0]), num_samples=10000,
                           time_step=0.1, random_type=odraws.RandomType.PSEUDO, seed=123)
  mc = np.mean(np.sum(np.diff(times) * paths[:, 1:, 1], axis=1))
  expected = (v0 - theta) * (1 - np.exp(-kappa * future_time)) / kappa + theta * future_time
  np.testing.assert_allclose(expected, mc, rtol=0.01)


def _heston_call(spot, strike, rate, expiry, kappa, theta, volvol, rho, v0):
  """Semi-analytic Heston call (Gil-Pelaez inversion of the log-spot characteristic
  function, 'little trap' form) -- an independent stand-in for the reference's
  `heston.approximations.european_option_price` (heston_model_test.py:363-375)."""
  def cf(u):
    d = np.sqrt((rho * volvol * 1j * u - kappa)**2 + volvol**2 * (1j * u + u * u))
    g = (kappa - rho * volvol * 1j * u - d) / (kappa - rho * volvol * 1j * u + d)
    e = np.exp(-d * expiry)
    c = kappa * theta / volvol**2 * ((kappa - rho * volvol * 1j * u - d) * expiry - 2 * np.log((1 - g * e) / (1 - g)))
    dd = (kappa - rho * volvol * 1j * u - d) / volvol**2 * (1 - e) / (1 - g * e)
    return np.exp(1j * u * (np.log(spot) + rate * expiry) + c + dd * v0)
  k = np.log(strike)
  p1 = 0.5 + integrate.quad(lambda u: (np.exp(-1j * u * k) * cf(u - 1j) / (1j * u * cf(-1j))).real, 1e-9, 200)[0] / np.pi
  p2 = 0.5 + integrate.quad(lambda u: (np.exp(-1j * u * k) * cf(u) / (1j * u)).real, 1e-9, 200)[0] / np.pi
  return spot * p1 - strike * np.exp(-rate * expiry) * p2


@pytest.mark.parametrize('mode', ['num_time_steps', 'time_step', 'times_grid', 'times_grid_and_draws'])
def test_heston_compare_monte_carlo_to_european_option(mode):
  # heston_model_test.py:270-378
  kappa, theta, volvol, rho = 0.3, 0.05, 0.02, 0.1
  maturity, log_spot, v0, strike, discounting = 1.0, 3.0, 0.05, 15, 0.5
  mean_reversion = omodels.PiecewiseConstantFunc([0.1, 0.2], [kappa, kappa, kappa], dtype=np.float64)
  kw = dict(num_samples=10000, random_type=odraws.RandomType.STATELESS_ANTITHETIC, seed=[1, 42])
  if mode == 'num_time_steps':
    kw['num_time_steps'] = 100
  else:
    kw['time_step'] = 0.01
  if mode.startswith('times_grid'):
    kw['times_grid'] = np.linspace(0.0, 1.0, 101)
  if mode == 'times_grid_and_draws':
    z = ophilox.stateless_normal([5000, 100, 2], [1, 42], np.float64)
    kw['num_samples'] = 1
    kw['normal_draws'] = np.concatenate([z, -z], axis=0)
  samples = oqe.sample_paths(mean_reversion, theta, volvol, rho, [maturity / 2, maturity],
                             np.array([log_spot, v0]), **kw)
  assert samples.shape == (10000, 2, 2)
  mc = np.exp(-discounting * maturity) * np.mean(
      np.maximum(np.exp(samples[:, -1, 0]) * np.exp(discounting * maturity) - strike, 0.0))
  want = _heston_call(np.exp(log_spot), strike, discounting, maturity, kappa, theta, volvol, rho, v0)
  assert abs(want - (np.exp(log_spot) - strike * np.exp(-discounting))) < 0.05      # deep in the money
  np.testing.assert_allclose(mc, want, atol=0.1, rtol=0.1)


# ---- batches of GBMs (`geometric_brownian_motion_test.py:284-420, 517-560`) ------------------------------
@pytest.mark.parametrize('batched_times', [False, True])
def test_univariate_sample_mean_constant_parameters_batched(batched_times):
  dtype = np.float64
  mu = np.array([[0.05], [0.06], [0.04], [0.03]], dtype=dtype)
  sigma = np.array([[0.05], [0.1], [0.15], [0.2]], dtype=dtype)
  times = (np.array([[0.1, 0.5, 1.0], [0.2, 0.4, 2.0], [0.3, 0.6, 5.0], [0.4, 0.9, 7.0]], dtype=dtype) if batched_times
           else np.array([0.1, 0.5, 1.0], dtype=dtype))
  x0 = np.array([[2.0], [10.0], [5.0], [25.0]], dtype=dtype)
  samples = omodels.gbm_exact_sample_paths(mu, sigma, times, initial_state=x0, num_samples=NUM_SAMPLES,
                                           random_type=odraws.RandomType.STATELESS, seed=[1234, 5], dtype=dtype)
  assert samples.shape == (4, NUM_SAMPLES, 3, 1)
  mean, var, se_mean, se_var = _log_moments(samples, NUM_SAMPLES)
  _within(mean, (mu - sigma**2 / 2) * times + np.log(x0), se_mean * NUM_STDERRS)
  _within(var, sigma**2 * times * np.ones((4, 1)), se_var * NUM_STDERRS)


def test_univariate_time_varying_drift_batched():
  # geometric_brownian_motion_test.py:517-560: batched piecewise drift, batched times, sigma = 0
  dtype = np.float64
  mu = omodels.PiecewiseConstantFunc(np.array([[0.0, 5.0, 10.0], [0.0, 7.0, 10.0]], dtype),
                                     np.array([[0.0, 0.0, 0.05, 0.05], [0.01, 0.01, 0.07, 0.07]], dtype), dtype=dtype)
  times = np.array([[0.0, 1.0, 5.0, 7.0, 10.0], [0.0, 1.5, 3.2, 4.8, 25.3]], dtype=dtype)
  samples = omodels.gbm_exact_sample_paths(mu, 0.0, times, initial_state=2.0, num_samples=1000,
                                           random_type=odraws.RandomType.STATELESS, seed=[1234, 5], dtype=dtype)
  assert samples.shape == (2, 1000, 5, 1)
  mean, var, _, _ = _log_moments(samples, 1000)
  expected = np.array([[0.0, 0.0, 0.0, 2.0 * 0.05, 5.0 * 0.05],
                       [0.0, 1.5 * 0.01, 3.2 * 0.01, 4.8 * 0.01, 7.0 * 0.01 + 18.3 * 0.07]]) + np.log(2.0)
  np.testing.assert_allclose(mean, expected, atol=1e-8)
  np.testing.assert_allclose(var, np.zeros((2, 5)), atol=1e-8)
