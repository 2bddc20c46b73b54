"""The reference's statistical Euler-sampler tests, run on the ORACLE (CPU).

`models/euler_sampling_test.py` holds no path values, only moments: these are
the checks that pin `oracle/euler.py` + `oracle/draws.py` (grid construction,
draw layout, step order, record plan) as a whole.  Same processes, sample
counts, generators, seeds and tolerances as the reference; the GPU side of the
same cases is `tests/test_gpu_generic_callables.py`.
"""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import philox as ophilox


def _wiener():
  return (lambda t, x: np.zeros_like(x)), (lambda t, x: np.ones_like(x)[..., None])


def _antithetic_draws(shape, seed, dtype=np.float64):
  # euler_sampling_test.py:100-107: stateless normals and their negatives
  z = ophilox.stateless_normal(shape, seed, dtype)
  return np.concatenate([z, -z], axis=-3)


@pytest.mark.parametrize('mode', ['time_step', 'num_time_steps', 'times_grid', 'times_grid_and_draws'])
def test_sample_paths_wiener(mode):
  # euler_sampling_test.py:71-133
  drift_fn, vol_fn = _wiener()
  times = np.array([0.1, 0.2, 0.3])
  num_samples = 10000
  kw = dict(num_samples=num_samples, random_type=odraws.RandomType.STATELESS_ANTITHETIC, seed=[1, 42])
  if mode == 'time_step':
    kw['time_step'] = 0.01
  elif mode == 'num_time_steps':
    kw['num_time_steps'] = 30
  else:
    kw['times_grid'] = np.linspace(0.0, 0.3, 31)
  if mode == 'times_grid_and_draws':
    kw['num_samples'] = 1
    kw['normal_draws'] = _antithetic_draws([5000, 30, 1], [1, 42])
  paths = oeuler.sample(1, drift_fn, vol_fn, times, dtype=np.float64, **kw)
  assert paths.shape == (num_samples, 3, 1)
  means = paths.mean(axis=0).reshape(-1)
  covars = np.cov(paths.reshape(num_samples, -1), rowvar=False)
  np.testing.assert_allclose(means, np.zeros(3), rtol=1e-2, atol=1e-2)
  np.testing.assert_allclose(covars, np.minimum(times[:, None], times[None, :]), rtol=1e-2, atol=1e-2)


def test_times_grid_long():
  # euler_sampling_test.py:135-172: a grid that runs past the last requested time
  drift_fn, vol_fn = _wiener()
  times = np.array([0.1, 0.2, 0.3])
  paths = oeuler.sample(1, drift_fn, vol_fn, times, num_samples=10000,
                        normal_draws=_antithetic_draws([5000, 32, 1], [1, 42]),
                        times_grid=np.linspace(0.0, 0.32, 33), seed=[1, 42], dtype=np.float64)
  assert paths.shape == (10000, 3, 1)
  np.testing.assert_allclose(paths.mean(axis=0).reshape(-1), np.zeros(3), rtol=1e-2, atol=1e-2)
  np.testing.assert_allclose(np.cov(paths.reshape(10000, -1), rowvar=False),
                             np.minimum(times[:, None], times[None, :]), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('use_batch,random_type,supply_normal_draws', [
    (False, 'STATELESS', False), (True, 'STATELESS', False), (True, 'STATELESS_ANTITHETIC', False),
    (True, 'STATELESS', True)])
def test_sample_paths_1d(use_batch, random_type, supply_normal_draws):
  # euler_sampling_test.py:174-294: dX = mu sqrt(t) dt + (a t + b) dW, E X_t = x0 + 2/3 mu t^1.5
  mu, a, b = 0.2, 0.4, 0.33
  drift_fn = lambda t, x: mu * np.sqrt(t) * np.ones_like(x)
  if use_batch:
    vol_fn = lambda t, x: (a * t + b) * np.ones([2, 1, 1, 1])
    x0 = np.array([[[0.1]], [[0.1]]])
  else:
    vol_fn = lambda t, x: (a * t + b) * np.ones([1, 1])
    x0 = np.array([0.1])
  times = np.array([0.0, 0.1, 0.21, 0.32, 0.43, 0.55])
  num_samples = 10000
  draws = _antithetic_draws([2, 5000, 55, 1], [1, 42]) if supply_normal_draws else None
  kw = dict(num_samples=num_samples, initial_state=x0, normal_draws=draws, time_step=0.01, seed=[1, 42],
            random_type=getattr(odraws.RandomType, random_type), dtype=np.float64)
  paths = oeuler.sample(1, drift_fn, vol_fn, times, **kw)
  paths_no_zero = oeuler.sample(1, drift_fn, vol_fn, times[1:], **kw)
  expected = x0 + (2.0 / 3.0) * mu * np.power(times, 1.5)
  if use_batch:
    assert paths.shape == (2, num_samples, 6, 1)
    np.testing.assert_allclose(paths.mean(axis=1).reshape(2, 1, 6), expected, rtol=1e-2, atol=1e-2)
  else:
    assert paths.shape == (num_samples, 6, 1)
    np.testing.assert_allclose(paths.mean(axis=0).reshape(-1), expected, rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(paths[:, 1:, :], paths_no_zero, rtol=1e-6, atol=1e-6)


_MU = np.array([0.2, 0.7])
_A = np.array([[0.4, 0.1], [0.3, 0.2]])
_B = np.array([[0.33, -0.03], [0.21, 0.5]])


@pytest.mark.parametrize('random_type,seed', [
    ('PSEUDO', 12134), ('STATELESS', [1, 2]), ('SOBOL', None), ('HALTON_RANDOMIZED', 12134)])
def test_sample_paths_2d(random_type, seed):
  # euler_sampling_test.py:296-361
  drift_fn = lambda t, x: _MU * np.sqrt(t) * np.ones_like(x)
  vol_fn = lambda t, x: (_A * t + _B) * np.ones([2, 2])
  times = np.array([0.1, 0.21, 0.32, 0.43, 0.55])
  x0 = np.array([0.1, -1.1])
  paths = oeuler.sample(2, drift_fn, vol_fn, times, num_samples=10000, initial_state=x0, time_step=0.01,
                        random_type=getattr(odraws.RandomType, random_type), seed=seed)
  assert paths.shape == (10000, 5, 2)
  expected = x0 + (2.0 / 3.0) * _MU * np.power(times[:, None], 1.5)
  np.testing.assert_allclose(paths.mean(axis=0), expected, rtol=1e-2, atol=1e-2)


def test_halton_sample_paths_2d():
  # euler_sampling_test.py:363-416
  drift_fn = lambda t, x: _MU * np.sqrt(t) * np.ones_like(x)
  vol_fn = lambda t, x: (_A * t + _B) * np.ones([2, 2])
  times = np.array([0.1, 0.21, 0.32])
  x0 = np.array([0.1, -1.1])
  paths = oeuler.sample(2, drift_fn, vol_fn, times, num_samples=10000, initial_state=x0, time_step=0.01,
                        random_type=odraws.RandomType.HALTON, seed=12134, skip=100, dtype=np.float64)
  assert paths.shape == (10000, 3, 2)
  expected = x0 + (2.0 / 3.0) * _MU * np.power(times[:, None], 1.5)
  np.testing.assert_allclose(paths.mean(axis=0), expected, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('random_type,seed', [('PSEUDO_ANTITHETIC', 12134), ('STATELESS_ANTITHETIC', [0, 12134])])
def test_antithetic_sample_paths_mean_2d(random_type, seed):
  # euler_sampling_test.py:418-482; the drift does not depend on the state
  drift_fn = lambda t, x: _MU * np.sqrt(t)
  vol_fn = lambda t, x: (_A * t + _B) * np.ones([2, 2])
  times = np.array([0.1, 0.21, 0.32, 0.43, 0.55])
  x0 = np.array([0.1, -1.1])
  paths = oeuler.sample(2, drift_fn, vol_fn, times, num_samples=5000, initial_state=x0, time_step=0.01,
                        random_type=getattr(odraws.RandomType, random_type), seed=seed)
  assert paths.shape == (5000, 5, 2)
  expected = x0 + (2.0 / 3.0) * _MU * np.power(times[:, None], 1.5)
  np.testing.assert_allclose(paths.mean(axis=0), expected, rtol=5e-3, atol=5e-3)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_sample_paths_dtypes(dtype):
  # euler_sampling_test.py:484-501
  drift_fn = lambda t, x: np.sqrt(t) * np.ones_like(x)
  vol_fn = lambda t, x: t * np.ones([1, 1], dtype=x.dtype)
  paths = oeuler.sample(1, drift_fn, vol_fn, [0.1, 0.2], num_samples=10, initial_state=[0.1],
                        time_step=0.01, seed=123, dtype=dtype)
  assert paths.dtype == dtype
