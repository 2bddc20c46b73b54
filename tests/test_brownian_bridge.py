"""Brownian-bridge barrier correction (SURVEY 8f-3;
`black_scholes/brownian_bridge.py:32-196`): the oracle and the stand-alone
mirror against the reference's documented values (CPU), the fused continuous-
barrier payoff against the oracle on the oracle's own paths (GPU)."""
import numpy as np
import pytest

from oracle import brownian_bridge as obb

X_START = np.asarray([[4.5, 4.5, 4.5], [4.5, 4.6, 4.7]])
X_END = np.asarray([[5.0, 4.9, 4.8], [4.8, 4.9, 5.0]])
VARIANCE = np.asarray([[0.1, 0.2, 0.1], [0.3, 0.1, 0.2]])


def test_oracle_double_barrier_documented_values():
  # brownian_bridge.py:55-72 (docstring example)
  got = obb.brownian_bridge_double(X_START, X_END, VARIANCE, 5.1, 4.4, n_cutoff=3)
  np.testing.assert_allclose(got, [[0.45842169, 0.21510919, 0.52704599],
                                   [0.09394963, 0.73302813, 0.22595022]], atol=1e-8)


def test_oracle_single_barrier_documented_values():
  # brownian_bridge.py:138-152 (docstring example)
  got = obb.brownian_bridge_single(X_START, X_END, VARIANCE, 5.1)
  np.testing.assert_allclose(got, [[0.69880579, 0.69880579, 0.97267628],
                                   [0.69880579, 0.86466472, 0.32967995]], atol=1e-8)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_mirror_matches_reference_test_construction(dtype):
  # brownian_bridge_test.py:35-127: numpy formula of the reference's own test
  import tff_b200 as tff
  x_start = np.asarray([[1.0, 1.1, 1.1], [1.05, 1.11, 1.11]], dtype=dtype)
  x_end = np.asarray([[2.0, 2.1, 2.8], [2.05, 2.11, 2.11]], dtype=dtype)
  variance = np.asarray([1.0, 1.0, 1.1], dtype=dtype)
  up, lo, n = 3.0, 0.5, 3

  def f(k):
    a = np.exp(-2 * k * (up - lo) * (k * (up - lo) + (x_end - x_start)) / variance)
    b = np.exp(-2 * (k * (up - lo) + x_start - up) * (k * (up - lo) + (x_end - up)) / variance)
    return a - b
  want = np.sum([f(k) for k in range(-n, n + 1)], axis=0)
  got = tff.black_scholes.brownian_bridge_double(
      x_start=x_start, x_end=x_end, variance=variance, dtype=dtype, upper_barrier=up,
      lower_barrier=lo, n_cutoff=n)
  assert tuple(got.shape) == want.shape
  np.testing.assert_allclose(got.cpu().numpy(), want, atol=1e-7)
  np.testing.assert_allclose(
      obb.brownian_bridge_double(x_start, x_end, variance, up, lo, n, dtype=dtype), want, atol=1e-7)
  barrier = 3.0
  want1 = 1 - np.exp(-2 * (x_start - barrier) * (x_end - barrier) / variance)
  got1 = tff.black_scholes.brownian_bridge_single(
      x_start=x_start, x_end=x_end, variance=variance, dtype=dtype, barrier=barrier)
  np.testing.assert_allclose(got1.cpu().numpy(), want1, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize('model', ['heston', 'log_gbm'])
def test_fused_continuous_barrier_matches_oracle(model):
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  from oracle import draws as odraws
  from oracle import euler as oeuler
  from oracle import grid as ogrid
  from oracle import models as omodels
  n, steps = 1 << 14, 50
  rt = tff.math.random.RandomType
  all_times, _, _ = ogrid.euler_grid([1.0], dtype=np.float64, num_time_steps=steps)
  dt = np.diff(all_times)
  if model == 'heston':
    m = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
    x0 = np.array([np.log(100.0), 0.04])
    od, ov = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, np.float64)
    kw = dict(random_type=rt.SOBOL, num_time_steps=steps)
    okw = dict(random_type=odraws.RandomType.SOBOL, num_time_steps=steps)
    dim, price = 2, lambda pay: m.price([1.0], pay, num_samples=n, initial_state=x0, **kw)
  else:
    r, sigma = 0.03, 0.25
    d, v = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
    proc = tff.models.GenericItoProcess(1, d, v, dtype=np.float64)
    x0 = np.array([np.log(100.0)])
    od = lambda t, x: (r - sigma**2 / 2) + 0 * x
    ov = lambda t, x: sigma * np.ones(x.shape + (1,))
    kw = dict(random_type=rt.STATELESS_ANTITHETIC, seed=[3, 9], num_time_steps=steps)
    okw = dict(random_type=odraws.RandomType.STATELESS_ANTITHETIC, seed=[3, 9], num_time_steps=steps)
    dim, price = 1, lambda pay: proc.price([1.0], pay, num_samples=n, initial_state=x0, **kw)
  up, dn = 125.0, 80.0
  pay = [engine.up_and_out_call(100.0, up, log_state=True, brownian_bridge=True),
         engine.up_and_out_call(100.0, up, log_state=True),
         engine.down_and_out_put(105.0, dn, log_state=True, brownian_bridge=True),
         engine.down_and_out_call(95.0, dn, log_state=True, brownian_bridge=True)]
  got = price(pay)
  paths = oeuler.sample(dim, od, ov, all_times[1:], times_grid=all_times, num_samples=n,
                        initial_state=x0, dtype=np.float64,
                        **{k: v for k, v in okw.items() if k != 'num_time_steps'})
  full = np.concatenate([np.broadcast_to(x0, (n, 1, dim)), paths], axis=1)      # [N, S+1, dim]
  xs, xe = full[:, :-1, 0], full[:, 1:, 0]
  var = (np.abs(full[:, :-1, 1]) if model == 'heston' else 0.25**2 * np.ones_like(xs)) * dt[None, :]

  def survive(level, upper):
    inner = (xs < level) & (xe < level) if upper else (xs > level) & (xe > level)
    p = obb.brownian_bridge_single(xs, xe, np.where(var > 0, var, 1.0), level)
    return np.prod(np.where(inner, np.where(var > 0, p, 1.0), 0.0), axis=1)
  st = np.exp(full[:, -1, 0])
  smax, smin = np.exp(full[:, :, 0]).max(axis=1), np.exp(full[:, :, 0]).min(axis=1)
  s_up, s_dn = survive(np.log(up), True), survive(np.log(dn), False)
  want = [np.where(smax > up, 0, np.maximum(st - 100, 0)) * s_up,
          np.where(smax > up, 0, np.maximum(st - 100, 0)),
          np.where(smin < dn, 0, np.maximum(105 - st, 0)) * s_dn,
          np.where(smin < dn, 0, np.maximum(st - 95, 0)) * s_dn]
  np.testing.assert_allclose(got, [w.mean() for w in want], rtol=1e-12)
  assert got[0] < got[1]          # continuous monitoring knocks out more paths
