"""`euler_sampling.sample` and the model classes' samplers: the HOST flow on the CPU.

`engine.Plan` is replaced by `tests/cpu_plan.CpuPlan` (numpy restatement of the kernels over
the tables and draw addressing the real plan would upload; the replacement is made by pytest's
`monkeypatch`, inside these tests only -- the package itself has no CPU path).  What runs is
everything the mirror does around a launch: argument handling and errors, time grids, the record
plan, recognising callables (model closures, probed plain callables), batches of processes and
their draw units, per-path initial states, supplied draws, antithetic layout, output shapes.  The
results are compared with the oracle's restatement of the reference, path by path; cases follow
`models/euler_sampling_test.py`, `generic_ito_process_test.py`, `heston_model_test.py`.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cpu_plan  # pylint: disable=g-import-not-at-top

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import heston_qe as oqe
from oracle import models as omodels
from oracle import philox as ophilox
import tff_b200 as tff
from tff_b200 import _tensor
from tff_b200 import engine

RT = odraws.RandomType


@pytest.fixture
def cpu_engine(monkeypatch):
  cpu_plan.install(monkeypatch)


def _rt(name):
  return getattr(tff.math.random.RandomType, name), getattr(RT, name)


MU, A, B = 0.2, 0.4, 0.33


@pytest.mark.parametrize('grid', [dict(time_step=0.01), dict(num_time_steps=30), dict(times_grid=np.linspace(0., 0.3, 31)),
                                  dict(times_grid=np.linspace(0., 0.32, 33))])
@pytest.mark.parametrize('random_type,seed', [('STATELESS_ANTITHETIC', [1, 42]), ('PSEUDO', 7), ('SOBOL', None)])
def test_wiener_process(cpu_engine, grid, random_type, seed):
  # euler_sampling_test.py:71-172: dX = dW under every way of specifying the grid
  prt, ort = _rt(random_type)
  times = np.array([0.1, 0.2, 0.3])
  got = tff.models.euler_sampling.sample(
      1, lambda t, x: torch.zeros_like(x), lambda t, x: torch.ones_like(x).unsqueeze(-1), times, num_samples=256,
      random_type=prt, seed=seed, dtype=np.float64, **grid)
  want = oeuler.sample(1, lambda t, x: np.zeros_like(x), lambda t, x: np.ones_like(x)[..., None], times,
                       num_samples=256, random_type=ort, seed=seed, dtype=np.float64, **grid)
  assert tuple(got.shape) == want.shape == (256, 3, 1)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-10, atol=1e-12)


def test_supplied_normal_draws(cpu_engine):
  # euler_sampling_test.py:100-107: antithetic draws handed in, num_samples taken from them
  times = np.array([0.1, 0.2, 0.3])
  z = ophilox.stateless_normal([128, 30, 1], [1, 42], np.float64)
  draws = np.concatenate([z, -z], axis=0)
  got = tff.models.euler_sampling.sample(
      1, lambda t, x: torch.zeros_like(x), lambda t, x: torch.ones_like(x).unsqueeze(-1), times, num_samples=1,
      normal_draws=torch.from_numpy(draws), times_grid=np.linspace(0., 0.3, 31), dtype=np.float64)
  want = oeuler.sample(1, lambda t, x: np.zeros_like(x), lambda t, x: np.ones_like(x)[..., None], times,
                       normal_draws=draws, times_grid=np.linspace(0., 0.3, 31), dtype=np.float64)
  assert tuple(got.shape) == (256, 3, 1)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-10, atol=1e-12)
  with pytest.raises(ValueError):          # euler_sampling.py:298-301
    tff.models.euler_sampling.sample(
        2, lambda t, x: torch.zeros_like(x), lambda t, x: torch.eye(2).expand(x.shape[0], 2, 2), times,
        normal_draws=torch.from_numpy(draws), times_grid=np.linspace(0., 0.3, 31), dtype=np.float64)


@pytest.mark.parametrize('use_batch,random_type', [(False, 'STATELESS'), (True, 'STATELESS'), (True, 'STATELESS_ANTITHETIC')])
def test_sample_paths_1d(cpu_engine, use_batch, random_type):
  # euler_sampling_test.py:174-294: dX = mu sqrt(t) dt + (a t + b) dW; a batch of two processes
  prt, ort = _rt(random_type)
  times = np.array([0.0, 0.1, 0.21, 0.32, 0.43, 0.55])
  # (a batch of initial states; plain callables whose VALUES carry a batch shape are outside what the
  # probing of `ProbedAffineSpec` accepts -- batched parameters go through the model classes)
  x0 = np.array([[[0.1]], [[0.3]]]) if use_batch else np.array([0.1])
  vol_t = lambda t, x: (A * t + B) * torch.ones([1, 1], dtype=torch.float64)
  vol_n = lambda t, x: (A * t + B) * np.ones([1, 1])
  kw = dict(num_samples=64, initial_state=x0, time_step=0.01, seed=[1, 42], dtype=np.float64)
  got = tff.models.euler_sampling.sample(1, lambda t, x: MU * torch.sqrt(t) * torch.ones_like(x), vol_t, times,
                                         random_type=prt, **kw)
  want = oeuler.sample(1, lambda t, x: MU * np.sqrt(t) * np.ones_like(x), vol_n, times, random_type=ort, **kw)
  assert tuple(got.shape) == want.shape == ((2, 64, 6, 1) if use_batch else (64, 6, 1))
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-9, atol=1e-11)
  if not use_batch:                        # the initial time is among the requested ones
    no_zero = tff.models.euler_sampling.sample(1, lambda t, x: MU * torch.sqrt(t) * torch.ones_like(x), vol_t,
                                               times[1:], random_type=prt, **kw)
    np.testing.assert_allclose(got.numpy()[:, 1:], no_zero.numpy(), rtol=1e-12)


MU2 = np.array([0.2, 0.7])
A2 = np.array([[0.4, 0.1], [0.3, 0.2]])
B2 = np.array([[0.33, -0.03], [0.21, 0.5]])


@pytest.mark.parametrize('random_type,seed,extra', [
    ('PSEUDO', 12134, {}), ('STATELESS', [1, 2], {}), ('SOBOL', None, {}), ('HALTON', None, {'skip': 100}),
    ('HALTON_RANDOMIZED', 12134, {}), ('PSEUDO_ANTITHETIC', 12134, {}), ('STATELESS_ANTITHETIC', [0, 12134], {})])
def test_sample_paths_2d(cpu_engine, random_type, seed, extra):
  # euler_sampling_test.py:296-482, plain callables of a 2-d process under every generator
  prt, ort = _rt(random_type)
  times = np.array([0.1, 0.21, 0.32, 0.43, 0.55])
  x0 = np.array([0.1, -1.1])
  kw = dict(num_samples=128, initial_state=x0, time_step=0.01, seed=seed, dtype=np.float64, **extra)
  got = tff.models.euler_sampling.sample(
      2, lambda t, x: torch.as_tensor(MU2) * torch.sqrt(t) * torch.ones_like(x),
      lambda t, x: (torch.as_tensor(A2) * t + torch.as_tensor(B2)) * torch.ones([2, 2], dtype=torch.float64), times,
      random_type=prt, **kw)
  want = oeuler.sample(2, lambda t, x: MU2 * np.sqrt(t) * np.ones_like(x), lambda t, x: (A2 * t + B2) * np.ones([2, 2]),
                       times, random_type=ort, **kw)
  assert tuple(got.shape) == want.shape == (128, 5, 2)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-9, atol=1e-11)


def test_per_path_initial_states(cpu_engine):
  # `initial_state` of shape [num_samples, dim] (`euler_sampling.py:357`)
  rs = np.random.RandomState(0)
  x0 = 100.0 * np.exp(0.1 * rs.standard_normal((32, 1)))
  gbm = tff.models.GeometricBrownianMotion(0.05, 0.3, dtype=np.float64)
  drift, vol = omodels.gbm_closures(0.05, 0.3, np.float64)
  kw = dict(num_samples=32, initial_state=x0, num_time_steps=10, seed=[3, 4], dtype=np.float64)
  got = tff.models.euler_sampling.sample(1, gbm.drift_fn(), gbm.volatility_fn(), [0.5, 1.0],
                                         random_type=tff.math.random.RandomType.STATELESS, **kw)
  want = oeuler.sample(1, drift, vol, [0.5, 1.0], random_type=RT.STATELESS, **kw)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)


def test_dtype_and_argument_errors(cpu_engine):
  # euler_sampling_test.py:484-570
  for dtype in (np.float32, np.float64):
    got = tff.models.euler_sampling.sample(
        1, lambda t, x: torch.sqrt(t) * torch.ones_like(x), lambda t, x: t * torch.ones([1, 1], dtype=x.dtype),
        [0.1, 0.2], num_samples=10, initial_state=[0.1], time_step=0.01, seed=123, dtype=dtype)
    assert got.numpy().dtype == dtype and tuple(got.shape) == (10, 2, 1)
  gbm = tff.models.GeometricBrownianMotion(0.05, 0.3, dtype=np.float64)
  with pytest.raises(ValueError):          # both time_step and num_time_steps (euler_sampling.py:254-258)
    tff.models.euler_sampling.sample(1, gbm.drift_fn(), gbm.volatility_fn(), [1.0], time_step=0.1, num_time_steps=5)
  with pytest.raises(ValueError):          # neither (euler_sampling.py:259-263)
    tff.models.euler_sampling.sample(1, gbm.drift_fn(), gbm.volatility_fn(), [1.0])
  with pytest.raises(NotImplementedError):  # a callable that is not affine in the state: no silent CPU fallback
    tff.models.euler_sampling.sample(1, lambda t, x: torch.sin(x), lambda t, x: torch.ones_like(x).unsqueeze(-1),
                                     [1.0], num_time_steps=4, seed=1, dtype=np.float64)


def test_model_classes_sample_paths(cpu_engine):
  rt_p, rt_o = _rt('STATELESS')
  # Heston: the Euler closures through GenericItoProcess.sample_paths, and the QE scheme of sample_paths
  pw = tff.math.piecewise.PiecewiseConstantFunc
  opw = omodels.PiecewiseConstantFunc
  args = ([0.5], [1.0, 1.1]), ([0.5], [0.04, 0.09]), ([0.3], [0.5, 0.8]), ([0.5], [-0.7, 0.6])
  heston = tff.models.HestonModel(*[pw(j, v, dtype=np.float64) for j, v in args], dtype=np.float64)
  oargs = [opw(j, v, dtype=np.float64) for j, v in args]
  x0 = np.array([np.log(100.0), 0.04])
  got = heston.sample_paths_euler([0.5, 1.0], x0, num_samples=64, time_step=0.05, random_type=rt_p, seed=[4, 2])
  drift, vol = omodels.heston_closures(*oargs, np.float64)
  want = oeuler.sample(2, drift, vol, [0.5, 1.0], num_samples=64, time_step=0.05, initial_state=x0,
                       random_type=rt_o, seed=[4, 2], dtype=np.float64)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-11, atol=1e-13)
  got = heston.sample_paths([0.5, 1.0], x0, num_samples=64, time_step=0.05, random_type=rt_p, seed=[4, 2])
  want = oqe.sample_paths(*oargs, [0.5, 1.0], x0, num_samples=64, time_step=0.05, random_type=rt_o, seed=[4, 2])
  assert tuple(got.shape) == want.shape == (64, 2, 2)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-10, atol=1e-12)

  # GBM: the exact log-normal sampler at the requested times only
  gbm = tff.models.GeometricBrownianMotion(pw([0.3], [0.05, 0.02], dtype=np.float64), 0.3, dtype=np.float64)
  got = gbm.sample_paths([0.25, 0.5, 1.0], initial_state=2.0, num_samples=64, random_type=rt_p, seed=[1234, 5])
  want = omodels.gbm_exact_sample_paths(opw([0.3], [0.05, 0.02], dtype=np.float64), 0.3, [0.25, 0.5, 1.0],
                                        initial_state=[2.0], num_samples=64, random_type=rt_o, seed=[1234, 5])
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)

  # multi-asset GBM (C4's model): Euler closures and the exact sampler
  means, vols, corr = np.array([0.05, 0.02, 0.03]), np.array([0.1, 0.2, 0.3]), np.array(
      [[1, 0.1, -0.2], [0.1, 1, 0.3], [-0.2, 0.3, 1]])
  mv = tff.models.MultivariateGeometricBrownianMotion(dim=3, means=means, volatilities=vols, corr_matrix=corr,
                                                      dtype=np.float64)
  x0 = np.array([1.0, 2.0, 3.0])
  got = mv.sample_paths([0.1, 0.5, 1.0], initial_state=x0, num_samples=64, random_type=rt_p, seed=[4, 2])
  want = omodels.mvgbm_exact_sample_paths(means, vols, corr, [0.1, 0.5, 1.0], initial_state=x0, num_samples=64,
                                          random_type=rt_o, seed=[4, 2])
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)
  got = tff.models.euler_sampling.sample(3, mv.drift_fn(), mv.volatility_fn(), [0.5, 1.0], num_samples=64,
                                         initial_state=x0, num_time_steps=8, random_type=rt_p, seed=[4, 2],
                                         dtype=np.float64)
  drift, vol = omodels.mvgbm_closures(means, vols, corr, np.float64)
  want = oeuler.sample(3, drift, vol, [0.5, 1.0], num_samples=64, initial_state=x0, num_time_steps=8,
                       random_type=rt_o, seed=[4, 2], dtype=np.float64)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)


# ---- the fused mode: `price` (nothing stored) ------------------------------------------------------
@pytest.fixture
def cpu_pricing(cpu_engine, monkeypatch):
  from tff_b200.models import euler_sampling
  monkeypatch.setattr(euler_sampling, '_CALLS', type(euler_sampling._CALLS)())


def test_heston_price_host_flow(cpu_pricing):
  # the public call `bench.py` times for C2: European + up-and-out call on the Euler scheme, Sobol,
  # barrier monitored on every grid point -- against payoffs evaluated on the oracle's paths
  heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
  x0 = np.array([np.log(100.0), 0.04])
  n, steps = 2048, 32
  payoffs = [engine.european_call(100.0, log_state=True), engine.up_and_out_call(100.0, 130.0, log_state=True),
             engine.european_put(95.0, log_state=True), engine.down_and_out_put(100.0, 85.0, log_state=True)]
  mean, stderr, bad = heston.price([1.0], payoffs, num_samples=n, initial_state=x0, num_time_steps=steps,
                                   random_type=tff.math.random.RandomType.SOBOL, return_stats=True)
  drift, vol = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, np.float64)
  paths, xmax, xmin = oeuler.sample(2, drift, vol, [1.0], num_samples=n, initial_state=x0, num_time_steps=steps,
                                    random_type=RT.SOBOL, dtype=np.float64, return_extrema=True)
  s, smax, smin = np.exp(paths[:, -1, 0]), np.exp(xmax), np.exp(xmin)
  want = np.stack([np.maximum(s - 100, 0), np.where(smax > 130, 0, np.maximum(s - 100, 0)), np.maximum(95 - s, 0),
                   np.where(smin < 85, 0, np.maximum(100 - s, 0))], -1)
  np.testing.assert_allclose(mean, want.mean(axis=0), rtol=1e-11)
  np.testing.assert_allclose(stderr, np.sqrt(np.maximum((want**2).mean(0) - want.mean(0)**2, 0) / n), rtol=1e-9)
  assert (bad == 0).all() and 0 < mean[1] < mean[0] and 0 < mean[3] < want[:, 2].mean() + 5


def test_c1_call_price_host_flow(cpu_pricing):
  # config C1: log-space GBM (additive noise -> TQF_MODEL_LINEAR_1F), PSEUDO_ANTITHETIC seed 42
  from tff_b200.models import closures
  r, sigma, spot, strike = 0.03, 0.2, 100.0, 100.0
  drift_fn, vol_fn = closures.affine_closures(r - 0.5 * sigma**2, 0.0, sigma)
  price = tff.models.euler_sampling.price(
      1, drift_fn, vol_fn, [1.0], [engine.european_call(strike, log_state=True, scale=np.exp(-r))],
      num_time_steps=100, num_samples=4096, initial_state=[np.log(spot)],
      random_type=tff.math.random.RandomType.PSEUDO_ANTITHETIC, seed=42, dtype=np.float64)
  paths = oeuler.sample(1, lambda t, x: (r - 0.5 * sigma**2) + 0 * x, lambda t, x: (sigma + 0 * x)[..., None], [1.0],
                        num_time_steps=100, num_samples=4096, initial_state=np.array([np.log(spot)]),
                        random_type=RT.PSEUDO_ANTITHETIC, seed=42, dtype=np.float64)
  want = np.exp(-r) * np.maximum(np.exp(paths[:, -1, 0]) - strike, 0).mean()
  np.testing.assert_allclose(price, [want], rtol=1e-11)
  from scipy.stats import norm
  d1 = (np.log(spot / strike) + r + 0.5 * sigma**2) / sigma
  assert abs(want - (spot * norm.cdf(d1) - strike * np.exp(-r) * norm.cdf(d1 - sigma))) < 0.5


def test_c3_swaption_price_host_flow(cpu_pricing):
  # config C3 through its public call: `hull_white.swaption_price(use_analytic_pricing=False)`; the reference's
  # case (swaption_test.py:85-125, 0.71632434 +- 1e-3) against the oracle's restatement of the same pricer
  from oracle import hull_white as ohw
  flat = lambda t: 0.01 * np.ones_like(np.asarray(t))
  kw = dict(expiries=np.array([1.0]), fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0]]),
            fixed_leg_daycount_fractions=0.25 * np.ones((1, 4)), fixed_leg_coupon=0.011 * np.ones((1, 4)),
            reference_rate_fn=flat, notional=100., mean_reversion=0.03, volatility=0.02, num_samples=1 << 15,
            time_step=0.1, seed=[4, 2], dtype=np.float64)
  got, stderr, bad = tff.models.hull_white.swaption_price(
      floating_leg_start_times=None, floating_leg_end_times=None, floating_leg_daycount_fractions=None,
      use_analytic_pricing=False, random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, return_stats=True, **kw)
  want = ohw.swaption_price_mc(random_type=RT.STATELESS_ANTITHETIC, **kw)
  assert got.shape == want.shape == (1,) and bad[0] == 0
  np.testing.assert_allclose(got, want, rtol=1e-11)
  assert abs(got[0] - 0.71632434) < 4 * stderr[0] + 2e-3


def test_pathwise_delta_and_vega_host_flow(cpu_pricing):
  # the notebook's second half: delta / vega of a call on log-space GBM carried as tangents and reduced in the
  # fused kernel (`TangentAffine1FModel`, `*_TANGENT` payoffs) -- against the oracle's tangent recursion
  from oracle import tangent as otangent
  from tff_b200.models import closures
  r, sigma, spot, strike, n, steps = 0.03, 0.2, 100.0, 100.0, 4096, 16
  # X = log S: a0 = r - sigma^2 / 2, b0 = sigma; d/dsigma: da0 = -sigma, db0 = 1
  drift_fn, vol_fn = closures.affine_tangent_closures(r - 0.5 * sigma**2, 0.0, sigma, 0.0, da0=-sigma, db=1.0)
  spec = drift_fn.tqf_spec
  payoffs = [engine.european_call(strike, log_state=True, scale=np.exp(-r)),
             engine.european_call_tangent(strike, spec.D_INITIAL, log_state=True, scale=np.exp(-r) / spot),
             engine.european_call_tangent(strike, spec.D_THETA, log_state=True, scale=np.exp(-r))]
  got = tff.models.euler_sampling.price(1, drift_fn, vol_fn, [1.0], payoffs, num_time_steps=steps, num_samples=n,
                                        initial_state=[np.log(spot)], random_type=tff.math.random.RandomType.SOBOL,
                                        dtype=np.float64)
  st = otangent.sample_with_tangents(r - 0.5 * sigma**2, 0.0, sigma, 0.0, -sigma, 0.0, 1.0, 0.0, [1.0],
                                     [np.log(spot)], n, random_type=RT.SOBOL, num_time_steps=steps)[:, -1]
  s = np.exp(st[:, 0])
  itm = s > strike
  want = [np.exp(-r) * np.maximum(s - strike, 0).mean(), np.exp(-r) / spot * np.where(itm, s * st[:, 1], 0).mean(),
          np.exp(-r) * np.where(itm, s * st[:, 2], 0).mean()]
  np.testing.assert_allclose(got, want, rtol=1e-10)
  from scipy.stats import norm
  d1 = (np.log(spot / strike) + r + 0.5 * sigma**2) / sigma
  assert abs(got[1] - norm.cdf(d1)) < 2e-2 and abs(got[2] - spot * norm.pdf(d1)) < 1.0     # Black-Scholes delta, vega


# ---- the Milstein sampler --------------------------------------------------------------------------
@pytest.mark.parametrize('random_type,dtype', [('STATELESS_ANTITHETIC', np.float64), ('SOBOL', np.float64),
                                               ('STATELESS', np.float32)])
def test_milstein_host_flow(cpu_engine, random_type, dtype):
  # milstein_sampling_test.py:165-230: dX = r X dt + sigma X dW, the correction sigma^2 X (dW^2 - dt) / 2 active;
  # the reference's `dim + 3 dim order` draw tensor is generated and its first column fed to the kernel
  from oracle import milstein as omilstein
  prt, ort = _rt(random_type)
  r, sigma = 0.5, 0.5
  times = np.array([0.0, 0.1, 0.21, 0.32, 0.43, 0.55], dtype=dtype)
  x0 = np.array([0.1], dtype=dtype)
  process = tff.models.GeometricBrownianMotion(r, sigma, dtype=dtype)
  kw = dict(num_samples=128, initial_state=x0, time_step=0.01, seed=[1, 42], skip=3)
  got = tff.models.milstein_sampling.sample(dim=1, drift_fn=process.drift_fn(), volatility_fn=process.volatility_fn(),
                                            times=times, random_type=prt, dtype=dtype, **kw)
  want = omilstein.sample(dim=1, drift_fn=lambda t, x: dtype(r) * x, volatility_fn=lambda t, x: (dtype(sigma) * x)[..., None],
                          grad_volatility_fn=lambda t, x: dtype(sigma) * np.ones(x.shape + (1,), dtype=dtype),
                          times=times, random_type=ort, dtype=dtype, **kw)
  assert tuple(got.shape) == want.shape == (128, 6, 1) and got.numpy().dtype == dtype
  if dtype == np.float64:
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)
  else:
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-5, atol=2e-7)


def test_milstein_nd_host_flow(cpu_engine):
  # state-independent volatility matrix: the gradient is zero, the Stratonovich terms vanish and the step is the
  # Euler kernel on the first `dim` columns of the `dim + 3 dim order` draw tensor (`milstein_sampling.py:481-595`)
  from oracle import milstein as omilstein
  prt, ort = _rt('STATELESS')
  times = np.array([0.1, 0.21, 0.32])
  x0 = np.array([0.1, -1.1])
  kw = dict(num_samples=64, initial_state=x0, time_step=0.01, seed=[1, 42])
  got = tff.models.milstein_sampling.sample(
      dim=2, drift_fn=lambda t, x: torch.as_tensor(MU2) * torch.sqrt(t) * torch.ones_like(x),
      volatility_fn=lambda t, x: (torch.as_tensor(A2) * t + torch.as_tensor(B2)) * torch.ones([2, 2], dtype=torch.float64),
      times=times, random_type=prt, dtype=np.float64, **kw)
  want = omilstein.sample(dim=2, drift_fn=lambda t, x: MU2 * np.sqrt(t) * np.ones_like(x),
                          volatility_fn=lambda t, x: np.broadcast_to(A2 * t + B2, x.shape + (2,)),
                          grad_volatility_fn=lambda t, x: [np.zeros(x.shape + (2,)) for _ in range(2)],
                          times=times, random_type=ort, dtype=np.float64, **kw)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-9, atol=1e-11)


# ---- batches of processes: one plan per batch element, draw units strided / offset -----------------
@pytest.mark.parametrize('random_type,seed,skip', [('SOBOL', None, 3), ('STATELESS', [4, 2], 0),
                                                   ('STATELESS_ANTITHETIC', [4, 2], 0), ('PSEUDO_ANTITHETIC', 9, 0)])
def test_batch_of_initial_states(cpu_engine, random_type, seed, skip):
  # batch_shape = initial_state.shape[:-2] (euler_sampling.py:251); draws laid out [steps] + batch + [N, dim], or
  # [N / 2] + batch for the antithetic types (models/utils.py:98-128)
  prt, ort = _rt(random_type)
  heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
  odrift, ovol = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, np.float64)
  x0 = np.array([[[np.log(100.0), 0.04]], [[np.log(90.0), 0.09]], [[np.log(120.0), 0.01]]])
  kw = dict(num_samples=64, initial_state=x0, seed=seed, skip=skip, num_time_steps=6, dtype=np.float64)
  got = tff.models.euler_sampling.sample(2, heston.drift_fn(), heston.volatility_fn(), [0.5, 1.0], random_type=prt, **kw)
  want = oeuler.sample(2, odrift, ovol, [0.5, 1.0], random_type=ort, **kw)
  assert tuple(got.shape) == want.shape == (3, 64, 2, 2)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-11, atol=1e-13)


def test_batch_of_gbm_parameters(cpu_engine):
  # batched GBM parameters of shape batch_shape + [1] (univariate_geometric_brownian_motion.py:66-80)
  from tff_b200.models import closures
  mean, vol = np.array([[0.01], [0.05]]), np.array([[0.1], [0.3]])
  drift, volf = closures.gbm_closures(mean, vol)
  x0 = np.array([[[1.0]], [[2.0]]])
  kw = dict(num_samples=70, initial_state=x0, seed=[1, 5], time_step=0.1, dtype=np.float64)
  got = tff.models.euler_sampling.sample(1, drift, volf, [1.0], random_type=tff.math.random.RandomType.STATELESS, **kw)
  want = oeuler.sample(1, lambda t, x: mean[:, None, :] * x, lambda t, x: (vol[:, None, :] * x)[..., None], [1.0],
                       random_type=RT.STATELESS, **kw)
  assert tuple(got.shape) == want.shape == (2, 70, 1, 1)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)


# ---- Hull-White samplers (exact OU step) -----------------------------------------------------------
def _flat_rate(t):
  return 0.01 * np.ones_like(np.asarray(t))


def test_hull_white_1f_sample_paths(cpu_engine):
  # vector_hull_white.py:641-781 for dim 1: piecewise-constant volatility (jumps enter the grid), a times_grid
  from oracle import hull_white as ohw
  prt, ort = _rt('STATELESS_ANTITHETIC')
  pw = tff.math.piecewise.PiecewiseConstantFunc([0.1, 0.5], [0.01, 0.02, 0.015], dtype=np.float64)
  opw = omodels.PiecewiseConstantFunc([0.1, 0.5], [0.01, 0.02, 0.015], dtype=np.float64)
  model = tff.models.hull_white.HullWhiteModel1F(0.03, pw, _flat_rate, dtype=np.float64)
  omodel = ohw.HullWhiteModel1F(0.03, opw, _flat_rate, np.float64)
  for kw in (dict(), dict(times_grid=np.linspace(0.0, 1.0, 11))):
    got = model.sample_paths([0.1, 0.5, 1.0], num_samples=64, random_type=prt, seed=[1, 2], **kw)
    want = omodel.sample_paths([0.1, 0.5, 1.0], 64, ort, seed=[1, 2], **kw)
    assert tuple(got.shape) == want.shape == (64, 3, 1)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-11, atol=1e-14)


def test_vector_hull_white_sample_paths(cpu_engine):
  # hull_white_test.py:223-270: two correlated factors, one piecewise-constant volatility each
  from oracle import hull_white as ohw
  prt, ort = _rt('STATELESS_ANTITHETIC')
  a, sigma = np.array([0.1, 0.05]), np.array([0.01, 0.02])
  vol = tff.math.piecewise.PiecewiseConstantFunc([[0.1, 0.2, 0.5], [0.1, 2.0, 3.0]],
                                                 [[0.01, 0.012, 0.009, 0.01], [0.02, 0.02, 0.02, 0.02]], dtype=np.float64)
  ovols = [omodels.PiecewiseConstantFunc([0.1, 0.2, 0.5], [0.01, 0.012, 0.009, 0.01], dtype=np.float64),
           omodels.PiecewiseConstantFunc([0.1, 2.0, 3.0], 4 * [0.02], dtype=np.float64)]
  flat2 = lambda t: 0.01 * np.ones(np.shape(t) + (2,)) if not isinstance(t, torch.Tensor) else 0.01 * torch.ones(
      tuple(t.shape) + (2,), dtype=t.dtype)
  model = tff.models.hull_white.VectorHullWhiteModel(2, a, vol, flat2, corr_matrix=[[1., 0.5], [0.5, 1.]],
                                                     dtype=np.float64)
  omodel = ohw.VectorHullWhiteModel(2, a, ovols, _flat_rate, [[1., 0.5], [0.5, 1.]])
  got = model.sample_paths([0.1, 0.5, 1.0], num_samples=64, random_type=prt, seed=[1, 2])
  want = omodel.sample_paths([0.1, 0.5, 1.0], 64, ort, seed=[1, 2])
  assert tuple(got.shape) == want.shape == (64, 3, 2)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize('mode', ['num_time_steps', 'time_step', 'times_grid', 'times_grid_and_draws'])
def test_heston_qe_grid_modes(cpu_engine, mode):
  # heston_model_test.py:270-378: the four ways `HestonModel.sample_paths` is told its grid, piecewise mean reversion
  prt, ort = _rt('STATELESS_ANTITHETIC')
  jumps, vals = [0.1, 0.2], [0.3, 0.3, 0.3]
  heston = tff.models.HestonModel(mean_reversion=tff.math.piecewise.PiecewiseConstantFunc(jumps, vals, dtype=np.float64),
                                  theta=0.05, volvol=0.02, rho=0.1, dtype=np.float64)
  okappa = omodels.PiecewiseConstantFunc(jumps, vals, dtype=np.float64)
  x0 = np.array([3.0, 0.05])
  kw, okw = dict(num_samples=200, random_type=prt, seed=[1, 42]), dict(num_samples=200, random_type=ort, seed=[1, 42])
  extra = {}
  if mode == 'num_time_steps':
    extra['num_time_steps'] = 100
  else:
    extra['time_step'] = 0.01
  if mode.startswith('times_grid'):
    extra['times_grid'] = np.linspace(0.0, 1.0, 101)
  if mode == 'times_grid_and_draws':
    z = ophilox.stateless_normal([100, 100, 2], [1, 42], np.float64)
    draws = np.concatenate([z, -z], axis=0)
    kw.update(num_samples=1, normal_draws=torch.from_numpy(draws))
    okw.update(num_samples=1, normal_draws=draws)
  got = heston.sample_paths(times=[0.5, 1.0], initial_state=x0, **kw, **extra)
  want = oqe.sample_paths(okappa, 0.05, 0.02, 0.1, [0.5, 1.0], x0, **okw, **extra)
  assert tuple(got.shape) == want.shape == (200, 2, 2)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-10, atol=1e-12)


# ---- a seeded sweep over grids: requested times on / off / near grid points, every grid mode -------
@pytest.mark.parametrize('case', range(40))
def test_grid_sweep(cpu_engine, case):
  rs = np.random.RandomState(1000 + case)
  dtype = np.float64 if case % 4 else np.float32
  k = rs.randint(1, 6)
  mode = ('time_step', 'num_time_steps', 'times_grid')[case % 3]
  if mode == 'time_step':
    step = dtype([0.05, 0.1, 0.125, 0.3][rs.randint(4)])
    # some requested times are exact multiples of the step, some within the de-duplication tolerance of one
    times = np.sort(np.concatenate([step * rs.randint(1, 12, size=k), rs.uniform(0.01, 1.2, size=rs.randint(0, 3))]))
    if case % 5 == 0:
      times[0] = times[0] * (1 + 1e-12)
    grid = dict(time_step=step)
  elif mode == 'num_time_steps':
    times = np.sort(rs.uniform(0.01, 2.0, size=k))
    grid = dict(num_time_steps=int(rs.randint(1, 20)))
  else:
    g = np.unique(np.concatenate([[0.0], np.round(rs.uniform(0.0, 2.0, size=rs.randint(3, 25)), 3)]))
    times = np.sort(rs.choice(g[1:], size=min(k, g.shape[0] - 1), replace=False)) if case % 2 else np.sort(
        rs.uniform(0.01, g[-1], size=k))
    grid = dict(times_grid=g.astype(dtype))
  if case % 7 == 0:
    times = np.concatenate([[0.0], times])          # the initial time among the requested ones
  times = times.astype(dtype)
  gbm = tff.models.GeometricBrownianMotion(0.05, 0.3, dtype=dtype)
  drift, vol = omodels.gbm_closures(0.05, 0.3, dtype)
  kw = dict(num_samples=8, initial_state=np.array([1.5], dtype=dtype), seed=[case, 7], dtype=dtype, **grid)
  want = oeuler.sample(1, drift, vol, times, random_type=RT.STATELESS, **kw)
  got = tff.models.euler_sampling.sample(1, gbm.drift_fn(), gbm.volatility_fn(), times,
                                         random_type=tff.math.random.RandomType.STATELESS, **kw)
  assert tuple(got.shape) == want.shape
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize('case', range(8))
def test_swaption_batch_sweep(cpu_pricing, case):
  # batches of swaptions with expiries off the time_step grid (on the grid the reference's own bookkeeping
  # breaks, see tests/test_hw_replay.py), payer / receiver, piecewise volatility in half of the cases
  from oracle import hull_white as ohw
  rs = np.random.RandomState(50 + case)
  b = rs.randint(1, 4)
  expiries = np.sort(np.round(rs.uniform(0.3, 3.0, size=b), 2)) + 0.003
  pay = expiries[:, None] + 0.25 * np.arange(1, 5)[None, :]
  is_payer = rs.rand(b) < 0.5
  if case % 2:
    vol = tff.math.piecewise.PiecewiseConstantFunc([0.5, 1.7], [0.01, 0.02, 0.015], dtype=np.float64)
    ovol = omodels.PiecewiseConstantFunc([0.5, 1.7], [0.01, 0.02, 0.015], dtype=np.float64)
  else:
    vol = ovol = 0.015
  kw = dict(expiries=expiries, fixed_leg_payment_times=pay, fixed_leg_daycount_fractions=0.25 * np.ones((b, 4)),
            fixed_leg_coupon=rs.uniform(0.005, 0.02) * np.ones((b, 4)), reference_rate_fn=_flat_rate, notional=100.,
            mean_reversion=0.03, is_payer_swaption=is_payer, num_samples=512, time_step=[0.1, 0.25][case % 2],
            seed=[case, 2], dtype=np.float64)
  got = tff.models.hull_white.swaption_price(
      floating_leg_start_times=None, floating_leg_end_times=None, floating_leg_daycount_fractions=None,
      use_analytic_pricing=False, random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, volatility=vol, **kw)
  want = ohw.swaption_price_mc(random_type=RT.STATELESS_ANTITHETIC, volatility=ovol, **kw)
  assert got.shape == want.shape == (b,)
  np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize('case', range(24))
def test_generator_and_initial_state_sweep(cpu_engine, case):
  # random type x (single | batch | per-path) initial state x skip x dtype on the Heston closures
  rs = np.random.RandomState(300 + case)
  name = ['PSEUDO', 'STATELESS', 'SOBOL', 'PSEUDO_ANTITHETIC', 'STATELESS_ANTITHETIC', 'HALTON'][case % 6]
  prt, ort = _rt(name)
  seed = {'PSEUDO': 11 + case, 'PSEUDO_ANTITHETIC': 11 + case, 'SOBOL': None, 'HALTON': None}.get(name, [case, 3])
  skip = int(rs.randint(0, 50)) if name in ('SOBOL', 'HALTON') else 0
  dtype = np.float32 if case % 8 == 7 else np.float64
  n = 32
  shape_kind = ('single', 'batch', 'per_path')[(case // 6) % 3]
  if name == 'HALTON' and shape_kind == 'batch':
    shape_kind = 'single'                   # (batched processes with HALTON draws are refused by the mirror)
  if shape_kind == 'single':
    x0 = np.array([np.log(100.0), 0.04])
  elif shape_kind == 'batch':
    x0 = np.stack([np.log(rs.uniform(80, 120, size=3)), rs.uniform(0.01, 0.09, size=3)], -1)[:, None, :]
  else:
    x0 = np.stack([np.log(rs.uniform(80, 120, size=n)), rs.uniform(0.01, 0.09, size=n)], -1)
  heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=dtype)
  odrift, ovol = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, dtype)
  kw = dict(num_samples=n, initial_state=x0.astype(dtype), seed=seed, skip=skip, num_time_steps=5, dtype=dtype)
  got = tff.models.euler_sampling.sample(2, heston.drift_fn(), heston.volatility_fn(), [0.4, 1.0], random_type=prt, **kw)
  want = oeuler.sample(2, odrift, ovol, [0.4, 1.0], random_type=ort, **kw)
  assert tuple(got.shape) == want.shape
  if dtype == np.float64:
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-11, atol=1e-13)
  else:
    np.testing.assert_allclose(got.numpy(), want, rtol=2e-5, atol=2e-6)


def test_heston_qe_price_host_flow(cpu_pricing):
  # the C2-QE bench line's public call: `HestonModel.price(scheme='qe')`, barrier monitored on every grid point
  heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
  x0 = np.array([np.log(100.0), 0.04])
  n, steps = 1024, 16
  payoffs = [engine.european_call(100.0, log_state=True), engine.up_and_out_call(100.0, 130.0, log_state=True)]
  mean, stderr, bad = heston.price([1.0], payoffs, num_samples=n, initial_state=x0, num_time_steps=steps,
                                   random_type=tff.math.random.RandomType.SOBOL, scheme='qe', return_stats=True)
  paths, xmax, _ = oqe.sample_paths(2.0, 0.04, 0.5, -0.7, [1.0], x0, num_samples=n, num_time_steps=steps,
                                    random_type=RT.SOBOL, return_extrema=True)
  s = np.exp(paths[:, -1, 0])
  want = np.stack([np.maximum(s - 100, 0), np.where(np.exp(xmax) > 130, 0, np.maximum(s - 100, 0))], -1)
  np.testing.assert_allclose(mean, want.mean(axis=0), rtol=1e-10)
  assert (bad == 0).all() and (stderr > 0).all()
  with pytest.raises(ValueError):
    heston.price([1.0], payoffs, num_samples=n, initial_state=x0, num_time_steps=steps, scheme='milstein')
  assert abs(heston.expected_total_variance(1.2, 0.3) - ((0.3 - 0.04) * (1 - np.exp(-2.0 * 1.2)) / 2.0 + 0.04 * 1.2)) < 1e-15


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_c4_basket_price_host_flow(cpu_pricing, dtype):
  # config C4 at a small size: correlated multi-asset GBM, Sobol, basket call (`component=-1`: mean over the assets)
  dim, n, steps = 8, 512, 12
  means, vols = np.full(dim, 0.03, dtype), np.linspace(0.1, 0.4, dim).astype(dtype)
  corr = (0.3 + 0.7 * np.eye(dim)).astype(dtype)
  mv = tff.models.MultivariateGeometricBrownianMotion(dim, means=means, volatilities=vols, corr_matrix=corr, dtype=dtype)
  x0 = 100.0 * np.ones(dim, dtype)
  got = tff.models.euler_sampling.price(dim, mv.drift_fn(), mv.volatility_fn(), np.array([1.0], dtype),
                                        [engine.european_call(100.0, component=-1)], num_time_steps=steps,
                                        num_samples=n, initial_state=x0, random_type=tff.math.random.RandomType.SOBOL,
                                        skip=5, dtype=dtype)
  d, v = omodels.mvgbm_closures(means, vols, corr, dtype)
  paths = oeuler.sample(dim, d, v, np.array([1.0], dtype), num_time_steps=steps, num_samples=n, initial_state=x0,
                        random_type=RT.SOBOL, skip=5, dtype=dtype)
  want = np.maximum(paths[:, 0, :].astype(np.float64).mean(axis=1) - 100.0, 0).mean()
  np.testing.assert_allclose(got, [want], rtol=1e-11 if dtype == np.float64 else 2e-5)


def test_c5_path_generation_host_flow(cpu_engine):
  # config C5's first half: log-space GBM Euler paths at the 50 exercise dates, stored exponentiated with their
  # column sums (what `least_square_mc(..., column_sums=)` consumes), STATELESS_ANTITHETIC -- the plan-level calls
  # `bench.py` and the README's multi-GPU example make
  from tff_b200 import distributed
  from tff_b200.models import closures, euler_sampling
  r, sigma, n = 0.06, 0.2, 256
  drift_fn, vol_fn = closures.affine_closures(r - 0.5 * sigma**2, 0.0, sigma)
  times = np.linspace(0.02, 1.0, 50)
  plans, record_slot, k, _ = euler_sampling._prepare(
      1, drift_fn, vol_fn, times, 0.01, None, n, [0.0], tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0,
      None, None, None, False, None, np.float64)
  plan = plans[0]
  lo, count = distributed.shard_units(plan.units)
  assert (lo, count) == (0, n // 2)
  paths, sums = plan.paths(record_slot, k, lo, count, exp_transform=True, column_sums=True)
  want = np.exp(oeuler.sample(1, lambda t, x: (r - 0.5 * sigma**2) + 0 * x, lambda t, x: (sigma + 0 * x)[..., None],
                              times, time_step=0.01, num_samples=n, initial_state=np.array([0.0]),
                              random_type=RT.STATELESS_ANTITHETIC, seed=[4, 2], dtype=np.float64))
  assert tuple(paths.shape) == want.shape == (n, 50, 1) and plan.num_steps == 100
  np.testing.assert_allclose(paths.numpy(), want, rtol=1e-12)
  np.testing.assert_allclose(sums.numpy(), want.sum(axis=0), rtol=1e-12)
  # the oracle's Longstaff-Schwartz price on those paths is the C5 number the GPU passes must reproduce
  from oracle import lsm as olsm
  df = np.exp(-r * times)
  price = olsm.least_square_mc(want, np.arange(50), olsm.make_basket_put_payoff([1.1]), olsm.make_polynomial_basis(3),
                               df, dtype=np.float64)
  assert price.shape == (1,) and 0.09 < price[0] < 0.2


@pytest.mark.parametrize('model', ['heston', 'log_gbm'])
def test_continuous_barrier_host_flow(cpu_pricing, model):
  # `brownian_bridge=True`: the knock-out payoffs weighted by the bridge's no-touch probability between grid points
  # (`black_scholes/brownian_bridge.py:118-196`); the same case as tests/test_brownian_bridge.py runs on the GPU
  from oracle import brownian_bridge as obb
  from oracle import grid as ogrid
  from tff_b200.models import closures
  n, steps = 1024, 20
  rt = tff.math.random.RandomType
  all_times, _, _ = ogrid.euler_grid([1.0], dtype=np.float64, num_time_steps=steps)
  dt = np.diff(all_times)
  if model == 'heston':
    m = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
    x0 = np.array([np.log(100.0), 0.04])
    od, ov = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, np.float64)
    okw = dict(random_type=RT.SOBOL)
    dim, price = 2, lambda pay: m.price([1.0], pay, num_samples=n, initial_state=x0, random_type=rt.SOBOL,
                                        num_time_steps=steps)
  else:
    r, sigma = 0.03, 0.25
    d, v = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
    proc = tff.models.GenericItoProcess(1, d, v, dtype=np.float64)
    x0 = np.array([np.log(100.0)])
    od = lambda t, x: (r - sigma**2 / 2) + 0 * x
    ov = lambda t, x: sigma * np.ones(x.shape + (1,))
    okw = dict(random_type=RT.STATELESS_ANTITHETIC, seed=[3, 9])
    dim, price = 1, lambda pay: proc.price([1.0], pay, num_samples=n, initial_state=x0,
                                           random_type=rt.STATELESS_ANTITHETIC, seed=[3, 9], num_time_steps=steps)
  up, dn = 125.0, 80.0
  pay = [engine.up_and_out_call(100.0, up, log_state=True, brownian_bridge=True),
         engine.up_and_out_call(100.0, up, log_state=True),
         engine.down_and_out_put(105.0, dn, log_state=True, brownian_bridge=True),
         engine.down_and_out_call(95.0, dn, log_state=True, brownian_bridge=True)]
  got = price(pay)
  paths = oeuler.sample(dim, od, ov, all_times[1:], times_grid=all_times, num_samples=n, initial_state=x0,
                        dtype=np.float64, **okw)
  full = np.concatenate([np.broadcast_to(x0, (n, 1, dim)), paths], axis=1)
  xs, xe = full[:, :-1, 0], full[:, 1:, 0]
  var = (np.abs(full[:, :-1, 1]) if model == 'heston' else 0.25**2 * np.ones_like(xs)) * dt[None, :]

  def survive(level, upper):
    inner = (xs < level) & (xe < level) if upper else (xs > level) & (xe > level)
    p = obb.brownian_bridge_single(xs, xe, np.where(var > 0, var, 1.0), level)
    return np.prod(np.where(inner, np.where(var > 0, p, 1.0), 0.0), axis=1)
  st = np.exp(full[:, -1, 0])
  smax, smin = np.exp(full[:, :, 0]).max(axis=1), np.exp(full[:, :, 0]).min(axis=1)
  s_up, s_dn = survive(np.log(up), True), survive(np.log(dn), False)
  want = [np.where(smax > up, 0, np.maximum(st - 100, 0)) * s_up, np.where(smax > up, 0, np.maximum(st - 100, 0)),
          np.where(smin < dn, 0, np.maximum(105 - st, 0)) * s_dn, np.where(smin < dn, 0, np.maximum(st - 95, 0)) * s_dn]
  np.testing.assert_allclose(got, [w.mean() for w in want], rtol=1e-10)
  assert got[0] < got[1]


# ---- batches of GBMs through the model class (`univariate_geometric_brownian_motion.py:66-80, 261-317`) --------
def _gbm_batch_case(name, dtype):
  pw, opw = tff.math.piecewise.PiecewiseConstantFunc, omodels.PiecewiseConstantFunc
  if name == 'constant':                       # geometric_brownian_motion_test.py:284-302
    mu = np.array([[0.05], [0.06], [0.04], [0.02]], dtype=dtype)
    sigma = np.array([[0.05], [0.1], [0.15], [0.2]], dtype=dtype)
    return mu, sigma, mu, sigma, np.array([0.1, 0.5, 1.0], dtype=dtype), np.array([[2.0], [10.0], [5.0], [25.0]], dtype)
  if name == 'batched_times':                  # :318-340
    mu = np.array([[0.05], [0.06], [0.04], [0.03]], dtype=dtype)
    sigma = np.array([[0.05], [0.1], [0.15], [0.2]], dtype=dtype)
    times = np.array([[0.1, 0.5, 1.0], [0.2, 0.4, 2.0], [0.3, 0.6, 5.0], [0.4, 0.9, 7.0]], dtype=dtype)
    return mu, sigma, mu, sigma, times, np.array([[2.0], [10.0], [5.0], [25.0]], dtype)
  if name == 'rank4':                          # :355-405
    rs = np.random.RandomState(3)
    mu = (0.3 * rs.uniform(size=(2, 3, 4, 1))).astype(dtype)
    sigma = (0.2 * rs.uniform(size=(2, 3, 4, 1))).astype(dtype)
    times = np.reshape(np.arange(1., 1. + (2 * 3 * 4 * 7), 1., dtype=dtype), (2, 3, 4, 7)) / 20
    return mu, sigma, mu, sigma, times, np.ones_like(mu) * 100.0
  # 'piecewise': batched piecewise drift and volatility, batched times, scalar initial state (:517-560, 637-700)
  jm, vm = np.array([[0.0, 5.0, 10.0], [0.0, 7.0, 10.0]], dtype), np.array([[0.0, 0.0, 0.05, 0.05], [0.01, 0.01, 0.07, 0.07]], dtype)
  js, vs = np.array([[0.0, 5.0, 10.0], [0.0, 7.0, 10.0]], dtype), np.array([[0.0, 0.2, 0.4, 0.6], [0.1, 0.1, 0.3, 0.3]], dtype)
  times = np.array([[0.0, 1.0, 5.0, 7.0, 10.0], [0.0, 1.5, 3.2, 4.8, 25.3]], dtype=dtype)
  return (pw(jm, vm, dtype=dtype), pw(js, vs, dtype=dtype), opw(jm, vm, dtype=dtype), opw(js, vs, dtype=dtype), times,
          np.asarray(2.0, dtype))


@pytest.mark.parametrize('name', ['constant', 'batched_times', 'rank4', 'piecewise'])
@pytest.mark.parametrize('supply_draws', [False, True])
def test_batched_gbm_sample_paths(cpu_engine, name, supply_draws):
  dtype = np.float64
  mu, sigma, omu, osigma, times, x0 = _gbm_batch_case(name, dtype)
  k = times.shape[-1]
  draws = ophilox.stateless_normal([64, k, 1], [4, 2], dtype) if supply_draws else None
  process = tff.models.GeometricBrownianMotion(mu, sigma, dtype=dtype)
  got = process.sample_paths(times=times, initial_state=x0, random_type=tff.math.random.RandomType.STATELESS,
                             num_samples=64, seed=[1234, 5],
                             normal_draws=None if draws is None else torch.from_numpy(draws))
  want = omodels.gbm_exact_sample_paths(omu, osigma, times, initial_state=x0, num_samples=64,
                                        random_type=RT.STATELESS, seed=[1234, 5], dtype=dtype, normal_draws=draws)
  batch = {'constant': (4,), 'batched_times': (4,), 'rank4': (2, 3, 4), 'piecewise': (2,)}[name]
  assert tuple(got.shape) == want.shape == batch + (64, k, 1)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)
  # every element of the batch runs on the SAME normals (the reference draws them without a batch shape)
  if name == 'constant':
    one = tff.models.GeometricBrownianMotion(float(mu[2, 0]), float(sigma[2, 0]), dtype=dtype).sample_paths(
        times=times, initial_state=x0[2], random_type=tff.math.random.RandomType.STATELESS, num_samples=64,
        seed=[1234, 5], normal_draws=None if draws is None else torch.from_numpy(draws))
    np.testing.assert_allclose(got.numpy()[2], one.numpy(), rtol=1e-13)


def test_batched_gbm_euler_closures(cpu_engine):
  # the class's closures carry the batched parameters into `euler_sampling.sample` (one plan per batch element,
  # draws laid out [steps] + batch + [N, dim])
  mu = np.array([[0.01], [0.05]])
  sigma = np.array([[0.1], [0.3]])
  process = tff.models.GeometricBrownianMotion(mu, sigma, dtype=np.float64)
  x0 = np.array([[[1.0]], [[2.0]]])
  kw = dict(num_samples=70, initial_state=x0, seed=[1, 5], time_step=0.1, dtype=np.float64)
  got = tff.models.euler_sampling.sample(1, process.drift_fn(), process.volatility_fn(), [1.0],
                                         random_type=tff.math.random.RandomType.STATELESS, **kw)
  want = oeuler.sample(1, lambda t, x: mu[:, None, :] * x, lambda t, x: (sigma[:, None, :] * x)[..., None], [1.0],
                       random_type=RT.STATELESS, **kw)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)


@pytest.mark.parametrize('batch_rank', [1, 2])
def test_generic_ito_process_batch_sample_paths_2d(cpu_engine, batch_rank):
  # generic_ito_process_test.py:147-211: a batch (rank 1 and 2) of initial states of a 2-d process given by plain
  # callables, PSEUDO draws with an integer seed, through `GenericItoProcess.sample_paths`
  process = tff.models.GenericItoProcess(
      dim=2, drift_fn=lambda t, x: torch.as_tensor(MU2) * torch.sqrt(t) * torch.ones_like(x),
      volatility_fn=lambda t, x: (torch.as_tensor(A2) * t + torch.as_tensor(B2)) * torch.ones(list(x.shape) + [2],
                                                                                             dtype=torch.float64),
      dtype=np.float64)
  times = np.array([0.1, 0.21, 0.32, 0.43, 0.55])
  x0 = np.array([0.1, -1.1]) * np.ones([2] * batch_rank + [1, 2]) + 0.01 * np.arange(2**batch_rank).reshape(
      [2] * batch_rank + [1, 1])
  got = process.sample_paths(times, num_samples=40, initial_state=x0, time_step=0.01, seed=12134)
  want = oeuler.sample(2, lambda t, x: MU2 * np.sqrt(t) * np.ones_like(x),
                       lambda t, x: (A2 * t + B2) * np.ones(x.shape + (2,)), times, num_samples=40, initial_state=x0,
                       time_step=0.01, seed=12134, random_type=RT.PSEUDO, dtype=np.float64)
  assert tuple(got.shape) == want.shape == tuple([2] * batch_rank + [40, 5, 2])
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize('case', range(16))
def test_heston_qe_sweep(cpu_engine, case):
  # `HestonModel.sample_paths`: its own grid (jumps and requested times merged into the uniform grid, duplicates kept,
  # zero-length steps skipped by `tolerance`, parameters taken at `t + min(dt) / 2`; heston_model.py:322-460, 575-639)
  rs = np.random.RandomState(700 + case)
  prt, ort = _rt(['STATELESS', 'SOBOL', 'STATELESS_ANTITHETIC', 'PSEUDO'][case % 4])
  seed = {0: [case, 1], 1: None, 2: [case, 1], 3: 40 + case}[case % 4]
  step = [0.05, 0.1, 0.125, 0.25][rs.randint(4)]

  def param(lo, hi):
    if rs.rand() < 0.5:
      return (float(rs.uniform(lo, hi)),) * 2
    n = rs.randint(1, 3)
    jumps = np.sort(np.where(rs.rand(n) < 0.5, step * rs.randint(1, 6, size=n), rs.uniform(0.05, 1.0, size=n)))
    vals = rs.uniform(lo, hi, size=n + 1)
    return (tff.math.piecewise.PiecewiseConstantFunc(jumps, vals, dtype=np.float64),
            omodels.PiecewiseConstantFunc(jumps, vals, dtype=np.float64))
  params = [param(0.5, 3.0), param(0.02, 0.09), param(0.2, 1.0), param(-0.8, 0.8)]
  times = np.sort(np.concatenate([step * rs.randint(1, 9, size=2), rs.uniform(0.05, 1.5, size=rs.randint(0, 3))]))
  x0 = np.array([np.log(100.0), 0.04])
  heston = tff.models.HestonModel(*[p[0] for p in params], dtype=np.float64)
  kw = dict(num_samples=32, seed=seed, time_step=step)
  got = heston.sample_paths(times, x0, random_type=prt, **kw)
  want = oqe.sample_paths(*[p[1] for p in params], times, x0, random_type=ort, **kw)
  assert tuple(got.shape) == want.shape == (32, times.shape[0], 2)
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize('case', range(10))
def test_hull_white_bond_option_sweep(cpu_pricing, case):
  # batches of calls / puts with random expiries (some on the uniform grid, some at 0), piecewise volatility
  from oracle import hull_white as ohw
  rs = np.random.RandomState(900 + case)
  step = [0.1, 0.25][case % 2]
  b = rs.randint(1, 5)
  expiries = np.where(rs.rand(b) < 0.4, step * rs.randint(0, 8, size=b), np.round(rs.uniform(0.1, 2.0, size=b), 3))
  maturities = expiries + rs.uniform(0.25, 5.0, size=b)
  strikes = np.exp(-0.01 * (maturities - expiries)) * rs.uniform(0.97, 1.03, size=b)
  is_call = rs.rand(b) < 0.5
  if case % 3:
    vol = tff.math.piecewise.PiecewiseConstantFunc([0.5, 1.7], [0.01, 0.02, 0.015], dtype=np.float64)
    ovol = omodels.PiecewiseConstantFunc([0.5, 1.7], [0.01, 0.02, 0.015], dtype=np.float64)
  else:
    vol = ovol = 0.015
  kw = dict(strikes=strikes, expiries=expiries, maturities=maturities, discount_rate_fn=_flat_rate, mean_reversion=0.03,
            is_call_options=is_call, num_samples=256, time_step=step, seed=[case, 9])
  got = tff.models.hull_white.bond_option_price(use_analytic_pricing=False, volatility=vol, dtype=np.float64,
                                                random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, **kw)
  want = ohw.bond_option_price_mc(volatility=ovol, random_type=RT.STATELESS_ANTITHETIC, **kw)
  assert got.shape == want.shape == (b,)
  np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13)
