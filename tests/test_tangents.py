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
