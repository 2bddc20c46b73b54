"""Pathwise tangents of the 1-d affine Euler scheme: the
forward-mode sensitivities the reference obtains with `watch_params`
(`models/euler_sampling.py:393-402, 467-510`) and the notebook's delta / vega
(`Monte_Carlo_Euler_Scheme.ipynb` cells 22-28).

CPU: the oracle's tangent recursion against central finite differences of the
oracle's own sampler.  GPU: the in-kernel tangents against the oracle (1e-12),
the fused tangent payoffs against the materialised oracle paths and against
Black-Scholes delta / vega (Monte-Carlo tolerance)."""
import math

import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import tangent as otangent

R, SIGMA, SPOT = 0.03, 0.1, 700.0


def _oracle_x(a0, a1, b0, b1, x0, times, **kw):
  return oeuler.sample(
      1, lambda t, x: a0 + a1 * x, lambda t, x: (b0 + b1 * x)[..., None], times,
      initial_state=np.array([x0]), dtype=np.float64, **kw)[..., 0]


@pytest.mark.parametrize('rt', ['STATELESS_ANTITHETIC', 'SOBOL'])
def test_oracle_tangents_equal_finite_differences(rt):
  kw = dict(num_samples=512, random_type=odraws.RandomType[rt], seed=[4, 2], num_time_steps=20)
  a0, a1, b0, b1 = 0.02, -0.3, 0.15, 0.25
  # theta enters all four coefficients: a0 = theta^2, a1 = -3 theta, b0 = theta + .05, b1 = 2.5 theta
  th = 0.1
  coef = lambda t: (t * t + 0.01, -3 * t, t + 0.05, 2.5 * t)
  d = (2 * th, -3.0, 1.0, 2.5)
  times = [0.5, 1.0]
  got = otangent.sample_with_tangents(*coef(th), *d, times, 1.3, **kw)
  np.testing.assert_array_equal(got[..., 0], _oracle_x(*coef(th), 1.3, times, **kw))
  h = 1e-6
  fd_x0 = (_oracle_x(*coef(th), 1.3 + h, times, **kw) - _oracle_x(*coef(th), 1.3 - h, times, **kw)) / (2 * h)
  fd_th = (_oracle_x(*coef(th + h), 1.3, times, **kw) - _oracle_x(*coef(th - h), 1.3, times, **kw)) / (2 * h)
  np.testing.assert_allclose(got[..., 1], fd_x0, rtol=1e-8, atol=1e-9)
  np.testing.assert_allclose(got[..., 2], fd_th, rtol=1e-7, atol=1e-8)


def _bs(spot, strike, sigma, r, t):
  d1 = (math.log(spot / strike) + (r + sigma**2 / 2) * t) / (sigma * math.sqrt(t))
  nd1 = 0.5 * (1 + math.erf(d1 / math.sqrt(2)))
  delta = nd1
  vega = spot * math.exp(-d1 * d1 / 2) / math.sqrt(2 * math.pi) * math.sqrt(t)
  return delta, vega


@pytest.mark.gpu
@pytest.mark.parametrize('rt', ['STATELESS_ANTITHETIC', 'SOBOL', 'STATELESS'])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_gpu_tangent_paths_match_oracle(rt, dtype):
  import tff_b200 as tff
  from tff_b200.models import closures
  vol_t = lambda t: 0.15 + 0.05 * np.asarray(t)              # time-dependent coefficient
  p = (0.02, -0.3, vol_t, 0.25, 0.2, -3.0, 1.0, 2.5)
  kw = dict(num_samples=3000, seed=[4, 2], time_step=0.03)
  drift, vol = closures.affine_tangent_closures(*p)
  got = tff.models.euler_sampling.sample(
      1, drift, vol, [0.4, 1.0], initial_state=np.array([1.3]),
      random_type=tff.math.random.RandomType[rt], dtype=dtype, **kw).cpu().numpy()
  want = otangent.sample_with_tangents(*p, [0.4, 1.0], 1.3, random_type=odraws.RandomType[rt],
                                       dtype=dtype, **kw)
  assert got.shape == want.shape == (3000, 2, 3) and got.dtype == dtype
  if dtype == np.float64:
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
  else:
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
def test_gpu_fused_delta_vega_notebook_setup():
  # log-space GBM of the notebook: X = log S, a0 = r - sigma^2/2, b0 = sigma, theta = sigma
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  T = engine.TangentAffineSpec1F
  n, steps, expiry = 1 << 18, 50, 1.0
  strikes = [600.0, 650.0, 680.0]
  drift, vol = closures.affine_tangent_closures(R - SIGMA**2 / 2, 0.0, SIGMA, 0.0,
                                                da0=-SIGMA, db=1.0)
  disc = math.exp(-R * expiry)
  payoffs = []
  for k in strikes:
    payoffs += [engine.european_call(k, log_state=True, scale=disc),
                engine.european_call_tangent(k, T.D_INITIAL, log_state=True, scale=disc / SPOT),
                engine.european_call_tangent(k, T.D_THETA, log_state=True, scale=disc)]
  kw = dict(num_samples=n, initial_state=np.array([math.log(SPOT)]), seed=[4, 2],
            num_time_steps=steps, dtype=np.float64)
  rt = tff.math.random.RandomType.STATELESS_ANTITHETIC
  got = tff.models.euler_sampling.price(1, drift, vol, [expiry], payoffs[:8], random_type=rt, **kw)
  got = np.concatenate([got, tff.models.euler_sampling.price(1, drift, vol, [expiry], payoffs[8:],
                                                             random_type=rt, **kw)])
  # (a) the same estimators on the oracle's materialised tangent paths
  o = otangent.sample_with_tangents(R - SIGMA**2 / 2, 0.0, SIGMA, 0.0, -SIGMA, 0.0, 1.0, 0.0,
                                    [expiry], math.log(SPOT), n,
                                    random_type=odraws.RandomType.STATELESS_ANTITHETIC,
                                    seed=[4, 2], num_time_steps=steps)[:, 0, :]
  s = np.exp(o[:, 0])
  want = []
  for k in strikes:
    itm = s > k
    want += [disc * np.maximum(s - k, 0).mean(), disc / SPOT * (itm * s * o[:, 1]).mean(),
             disc * (itm * s * o[:, 2]).mean()]
  np.testing.assert_allclose(got, want, rtol=1e-11)
  # (b) Black-Scholes delta / vega (the notebook reports 1.7e-3 / 7.8e-2 at 200k paths)
  for i, k in enumerate(strikes):
    delta, vega = _bs(SPOT, k, SIGMA, R, expiry)
    assert abs(got[3 * i + 1] - delta) < 5e-3 * delta
    assert abs(got[3 * i + 2] - vega) < 0.1 * vega


# ------------------------------------------------------------------ Heston ----
HESTON = dict(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7)
H_X0 = np.array([np.log(100.0), 0.04])


def _oracle_heston(params, x0, times, **kw):
  from oracle import models as omodels
  d, v = omodels.heston_closures(params['mean_reversion'], params['theta'], params['volvol'],
                                 params['rho'], np.float64)
  return oeuler.sample(2, d, v, times, initial_state=np.asarray(x0, dtype=np.float64),
                       dtype=np.float64, **kw)


@pytest.mark.parametrize('rt', ['STATELESS_ANTITHETIC', 'SOBOL'])
def test_oracle_heston_tangents_equal_finite_differences(rt):
  kw = dict(num_samples=256, random_type=odraws.RandomType[rt], seed=[4, 2], num_time_steps=24)
  times = [0.5, 1.0]
  zero = dict(d_mean_reversion=0.0, d_theta=0.0, d_volvol=0.0, d_rho=0.0, d_initial_state=(0.0, 0.0))
  base = _oracle_heston(HESTON, H_X0, times, **kw)
  h = 1e-6
  cases = [('mean_reversion', 'd_mean_reversion'), ('theta', 'd_theta'), ('volvol', 'd_volvol'),
           ('rho', 'd_rho'), ('x0', None), ('v0', None)]
  for name, dname in cases:
    d = dict(zero)
    if name == 'x0':
      d['d_initial_state'] = (1.0, 0.0)
      up, dn = (_oracle_heston(HESTON, H_X0 + s * np.array([h, 0.0]), times, **kw) for s in (1, -1))
    elif name == 'v0':
      d['d_initial_state'] = (0.0, 1.0)
      up, dn = (_oracle_heston(HESTON, H_X0 + s * np.array([0.0, h]), times, **kw) for s in (1, -1))
    else:
      d[dname] = 1.0
      up, dn = (_oracle_heston(dict(HESTON, **{name: HESTON[name] + s * h}), H_X0, times, **kw)
                for s in (1, -1))
    got = otangent.heston_with_tangents(**HESTON, **d, times=times, initial_state=H_X0, **kw)
    np.testing.assert_array_equal(got[..., :2], base)          # the path itself is untouched
    fd = (up - dn) / (2 * h)
    # paths whose variance crosses zero inside the step have a kink in |V|: compare the bulk
    ok = np.abs(got[..., 2:] - fd) <= 1e-5 + 1e-5 * np.abs(fd)
    assert ok.mean() > 0.97, (name, ok.mean())
    np.testing.assert_allclose(np.median(got[..., 2:], axis=0), np.median(fd, axis=0), rtol=1e-5,
                               atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize('rt', ['STATELESS_ANTITHETIC', 'SOBOL', 'STATELESS'])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_gpu_heston_tangent_paths_match_oracle(rt, dtype):
  import tff_b200 as tff
  from tff_b200.models import closures
  d = dict(d_mean_reversion=0.3, d_theta=-0.2, d_volvol=1.0, d_rho=0.5, d_initial_state=(0.25, 1.0))
  kw = dict(num_samples=3000, seed=[4, 2], time_step=0.04)
  drift, vol = closures.heston_tangent_closures(**HESTON, **d)
  got = tff.models.euler_sampling.sample(
      2, drift, vol, [0.4, 1.0], initial_state=H_X0.astype(dtype),
      random_type=tff.math.random.RandomType[rt], dtype=dtype, **kw).cpu().numpy()
  want = otangent.heston_with_tangents(**HESTON, **d, times=[0.4, 1.0], initial_state=H_X0,
                                       random_type=odraws.RandomType[rt], dtype=dtype, **kw)
  assert got.shape == want.shape == (3000, 2, 4) and got.dtype == dtype
  if dtype == np.float64:
    np.testing.assert_allclose(got[..., :2], want[..., :2], rtol=1e-12, atol=1e-14)
    # the tangents carry 1 / (2 sqrt|V|): rounding differences of V at the 1e-16 level are
    # amplified where a path's variance comes close to zero
    np.testing.assert_allclose(got[..., 2:], want[..., 2:], rtol=1e-8, atol=1e-9)
  else:
    np.testing.assert_allclose(got[..., :2], want[..., :2], rtol=1e-5, atol=5e-6)
    ok = np.abs(got - want) <= 2e-4 * np.maximum(1.0, np.abs(want))
    assert ok.mean() > 0.995          # float32 tangents amplify the rounding of 1 / sqrt|V| near 0


@pytest.mark.gpu
def test_gpu_heston_fused_vega_and_delta():
  # d price / d V_0 and d price / d S_0 of a call, fused, against the same estimators on the
  # oracle's materialised tangent paths and against a finite difference of fused prices
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  T = engine.TangentHestonSpec
  n, steps, strike = 1 << 17, 50, 100.0
  rt = tff.math.random.RandomType.SOBOL
  kw = dict(num_samples=n, initial_state=H_X0, num_time_steps=steps, dtype=np.float64)
  # vol-of-vol 0.3 satisfies the Feller condition: with 0.5 the Euler variance hits zero on
  # many paths, where d sqrt|V| / dV is unbounded -- the pathwise estimator is then heavy-tailed
  # and sits 3% away from a finite difference at 2^17 paths
  HESTON = dict(mean_reversion=2.0, theta=0.04, volvol=0.3, rho=-0.7)
  out = {}
  for name, d0 in (('vega_v0', (0.0, 1.0)), ('delta_log', (1.0, 0.0))):
    drift, vol = closures.heston_tangent_closures(**HESTON, d_initial_state=d0)
    pay = [engine.european_call(strike, log_state=True),
           engine.european_call_tangent(strike, T.D_X, log_state=True)]
    got = tff.models.euler_sampling.price(2, drift, vol, [1.0], pay, random_type=rt, **kw)
    o = otangent.heston_with_tangents(**HESTON, d_mean_reversion=0.0, d_theta=0.0, d_volvol=0.0,
                                      d_rho=0.0, d_initial_state=d0, times=[1.0], initial_state=H_X0,
                                      num_samples=n, random_type=odraws.RandomType.SOBOL,
                                      num_time_steps=steps)[:, 0, :]
    s = np.exp(o[:, 0])
    want = [np.maximum(s - strike, 0).mean(), ((s > strike) * s * o[:, 2]).mean()]
    np.testing.assert_allclose(got, want, rtol=1e-10)
    out[name] = got
  # finite difference of the fused price in V_0 (same Sobol points): the pathwise vega
  model = tff.models.HestonModel(dtype=np.float64, **HESTON)
  h = 1e-4
  up, dn = (model.price([1.0], [engine.european_call(strike, log_state=True)], num_samples=n,
                        initial_state=H_X0 + s * np.array([0.0, h]), random_type=rt,
                        num_time_steps=steps)[0] for s in (1, -1))
  np.testing.assert_allclose(out['vega_v0'][1], (up - dn) / (2 * h), rtol=1e-2)
  assert 0.4 < out['delta_log'][1] / 100.0 < 0.9          # d price / d S_0 = (d price / d log S_0) / S_0
