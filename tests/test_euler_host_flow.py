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
from oracle import halton as ohalton
from oracle import heston_qe as oqe
from oracle import models as omodels
from oracle import philox as ophilox
import tff_b200 as tff
from tff_b200 import _tensor
from tff_b200 import engine

RT = odraws.RandomType


@pytest.fixture
def cpu_engine(monkeypatch):
  monkeypatch.setattr(engine, 'Plan', cpu_plan.CpuPlan)
  monkeypatch.setattr(engine, 'cached_plan', lambda *a, **k: cpu_plan.CpuPlan(*a, **k))
  monkeypatch.setattr(_tensor, 'device', lambda: torch.device('cpu'))
  from tff_b200.math.random import halton

  def halton_normal(dim, n, skip=0, dtype=None, randomized=False, seed=None, randomization_params=None):
    u = ohalton.sample(dim, sequence_indices=np.arange(skip, skip + n), dtype=dtype, randomized=randomized, seed=seed)
    return torch.from_numpy(odraws._erfinv_times_sqrt2(u, dtype))
  monkeypatch.setattr(halton, 'sample_normal', halton_normal)


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
