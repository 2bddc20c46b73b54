"""One-dimensional Milstein sampler (`models/milstein_sampling.py`,
tests after `milstein_sampling_test.py:37-230`): the kernel against the oracle on
the same draws, plus the reference's own statistical checks."""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import milstein as omilstein

pytestmark = pytest.mark.gpu


def _tff():
  import tff_b200 as tff
  return tff


@pytest.mark.parametrize('use_time_step', [True, False])
def test_sample_paths_wiener(use_time_step):
  # milstein_sampling_test.py:37-107 -- plain Python callables, as in the reference
  tff = _tff()
  import torch
  times = np.array([0.1, 0.2, 0.3])
  n = 5000
  kw = dict(time_step=0.02) if use_time_step else dict(num_time_steps=15)
  paths = tff.models.milstein_sampling.sample(
      dim=1, drift_fn=lambda _, x: torch.zeros_like(x),
      volatility_fn=lambda _, x: torch.ones_like(x).unsqueeze(-1), times=times, num_samples=n,
      seed=[1, 42], random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, **kw).cpu().numpy()
  assert paths.shape == (n, 3, 1)
  np.testing.assert_allclose(paths.mean(axis=0).reshape(-1), np.zeros(3), rtol=1e-2, atol=1e-2)
  np.testing.assert_allclose(np.cov(paths.reshape(n, -1), rowvar=False),
                             np.minimum(times.reshape(-1, 1), times.reshape(1, -1)),
                             rtol=1e-2, atol=1e-2)
  want = omilstein.sample(
      dim=1, drift_fn=lambda t, x: np.zeros_like(x), volatility_fn=lambda t, x: np.ones(x.shape + (1,)),
      grad_volatility_fn=lambda t, x: np.zeros(x.shape + (1,)), times=times, num_samples=n,
      seed=[1, 42], random_type=odraws.RandomType.STATELESS_ANTITHETIC, **kw)
  np.testing.assert_allclose(paths, want, rtol=1e-12, atol=1e-14)


def test_sample_paths_1d_time_dependent():
  # dX = mu sqrt(t) dt + (a t + b) dW, milstein_sampling_test.py:109-163
  tff = _tff()
  import torch
  mu, a, b = 0.2, 0.4, 0.33
  times = np.array([0.0, 0.1, 0.21, 0.32, 0.43, 0.55])
  n, x0 = 10000, np.array([0.1])
  kw = dict(dim=1, drift_fn=lambda t, x: mu * torch.sqrt(t) * torch.ones_like(x),
            volatility_fn=lambda t, x: (a * t + b) * torch.ones([1, 1], dtype=t.dtype),
            num_samples=n, initial_state=x0,
            random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, time_step=0.01, seed=[1, 42])
  paths = tff.models.milstein_sampling.sample(times=times, **kw).cpu().numpy()
  paths_no_zero = tff.models.milstein_sampling.sample(times=times[1:], **kw).cpu().numpy()
  assert paths.shape == (n, 6, 1)
  np.testing.assert_allclose(paths.mean(axis=0).reshape(-1),
                             x0 + (2.0 / 3.0) * mu * np.power(times, 1.5), rtol=1e-2, atol=1e-2)
  np.testing.assert_allclose(paths[:, 1:, :], paths_no_zero)
  want = omilstein.sample(
      dim=1, drift_fn=lambda t, x: mu * np.sqrt(t) * np.ones_like(x),
      volatility_fn=lambda t, x: (a * t + b) * np.ones(x.shape + (1,)),
      grad_volatility_fn=lambda t, x: np.zeros(x.shape + (1,)), times=times, num_samples=n,
      initial_state=x0, random_type=odraws.RandomType.STATELESS_ANTITHETIC, time_step=0.01,
      seed=[1, 42])
  np.testing.assert_allclose(paths, want, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize('rt', ['STATELESS_ANTITHETIC', 'SOBOL', 'STATELESS'])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_sample_bsm_state_dependent_volatility(rt, dtype):
  # dX = r X dt + sigma X dW (not in log space), milstein_sampling_test.py:165-230:
  # the Milstein correction sigma^2 X (dW^2 - dt) / 2 is active here
  tff = _tff()
  r, sigma = 0.5, 0.5
  times = np.array([0.0, 0.1, 0.21, 0.32, 0.43, 0.55], dtype=dtype)
  n, x0 = 10000, np.array([0.1], dtype=dtype)
  process = tff.models.GeometricBrownianMotion(r, sigma, dtype=dtype)
  kw = dict(num_samples=n, initial_state=x0, time_step=0.01, seed=[1, 42], skip=3)
  paths = tff.models.milstein_sampling.sample(
      dim=1, drift_fn=process.drift_fn(), volatility_fn=process.volatility_fn(), times=times,
      random_type=tff.math.random.RandomType[rt], dtype=dtype, **kw).cpu().numpy()
  assert paths.shape == (n, 6, 1) and paths.dtype == dtype
  want = omilstein.sample(
      dim=1, drift_fn=lambda t, x: dtype(r) * x, volatility_fn=lambda t, x: (dtype(sigma) * x)[..., None],
      grad_volatility_fn=lambda t, x: dtype(sigma) * np.ones(x.shape + (1,), dtype=dtype), times=times,
      random_type=odraws.RandomType[rt], dtype=dtype, **kw)
  if dtype == np.float64:
    np.testing.assert_allclose(paths, want, rtol=1e-12)
  else:
    np.testing.assert_allclose(paths, want, rtol=1e-5, atol=2e-7)
  # E[X_t] = x0 exp(r t)
  np.testing.assert_allclose(paths.mean(axis=0).reshape(-1), x0 * np.exp(r * times),
                             rtol=2e-2, atol=1e-3)


def test_milstein_argument_errors():
  tff = _tff()
  from tff_b200.models import closures
  drift, vol = closures.gbm_closures(0.1, 0.2)
  with pytest.raises(NotImplementedError):
    tff.models.milstein_sampling.sample(dim=2, drift_fn=drift, volatility_fn=vol, times=[1.0],
                                        time_step=0.1)
  with pytest.raises(ValueError):
    tff.models.milstein_sampling.sample(dim=1, drift_fn=drift, volatility_fn=vol, times=[1.0])
  with pytest.raises(ValueError):
    tff.models.milstein_sampling.sample(dim=1, drift_fn=drift, volatility_fn=vol, times=[1.0],
                                        time_step=0.1, num_time_steps=10)
