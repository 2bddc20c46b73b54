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


def _unpack27(row, K):
  lhs = np.zeros((K, K))
  idx = 0
  for i in range(6):
    for j in range(i, 6):
      if i < K and j < K:
        lhs[i, j] = lhs[j, i] = row[idx]
      idx += 1
  return lhs, row[21:21 + K]


def test_c5_american_put_148_steps_50_dates_cubic_basis():
  # lsm.py:145-183 at the C5 shape: time_step 0.01 (148 Euler steps), 50 exercise dates,
  # STATELESS_ANTITHETIC seed [4, 2], cubic basis; N = 2^17.  The device path is the
  # persistent single-launch backward induction.
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  from tff_b200.models import utils
  from oracle import lsm as olsm
  lsm = tff.models.longstaff_schwartz
  n, r, sigma = 1 << 17, 0.1, 1.0
  times = np.linspace(0.0, 1.0, 50)
  drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.01), dtype=np.float64)
  nsteps, record_slot = engine.record_plan(mask, 50)
  assert nsteps == 148
  rng = engine.RngSpec(tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0)
  plan = engine.Plan(closures.resolve_spec(drift, vol), all_times, nsteps, np.array([0.0]), rng, n,
                     np.float64)
  try:
    paths, csums = plan.paths(record_slot, 50, exp_transform=True, column_sums=True)
  finally:
    plan.close()
  df = np.exp(-r * times)
  diag = {}
  got = lsm.least_square_mc(paths, np.arange(50), lsm.make_basket_put_payoff([1.1], dtype=np.float64),
                            lsm.make_polynomial_basis(3), discount_factors=df, dtype=np.float64,
                            column_sums=csums, diagnostics=diag)
  assert diag['route'] == 'persistent'

  jobs = [dict(lo=lo, hi=hi, n=n, r=r, sigma=sigma, times=times.tolist(), time_step=0.01,
               random_type='STATELESS_ANTITHETIC', seed=[4, 2]) for lo, hi in chunked.slices(n // 2, 1 << 12)]
  parts = chunked.run('gbm_log_paths', jobs)
  half = [p.shape[0] // 2 for p in parts]
  olog = np.concatenate([p[:h] for p, h in zip(parts, half)] + [p[h:] for p, h in zip(parts, half)])
  opaths = np.exp(olog)                                            # [N, 50, 1]
  np.testing.assert_allclose(paths.cpu().numpy(), opaths, rtol=1e-12)   # 148-step Euler paths
  odiag = {}
  want = olsm.least_square_mc(opaths, np.arange(50), olsm.make_basket_put_payoff([1.1]),
                              olsm.make_polynomial_basis(3), df, dtype=np.float64, diagnostics=odiag)
  # exercise decisions: paths whose final cashflow differs (a decision flipped on the way)
  w_got = diag['w'].cpu().numpy()
  w_want = odiag['w'][:, 0]
  flips = int(np.sum(np.abs(w_got - w_want) > 1e-9 * np.maximum(np.abs(w_want), 1e-3)))
  assert flips <= 2, flips
  # per-date normal equations X'X, X'y of the 49 regressions
  worst = 0.0
  for j in range(49):
    e = 49 - j
    lhs, rhs = _unpack27(diag['sums'][j], 4)
    for g, w in ((lhs, odiag['lhs'][e][0]), (rhs, odiag['rhs'][e][0])):
      err = np.abs(g - w).max() / np.abs(w).max()
      worst = max(worst, err)
  tol = 1e-13 if flips == 0 else 1e-4
  assert worst < tol, (worst, flips)
  np.testing.assert_allclose(got, want, rtol=1e-12 if flips == 0 else 1e-6)
  assert abs(float(got[0]) - 0.397) < 5e-3                          # docstring value at N = 1e5
