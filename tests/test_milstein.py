"""Milstein sampler (`models/milstein_sampling.py`, tests after
`milstein_sampling_test.py:37-323`): the kernel against the oracle on the same
draws, plus the reference's own statistical checks.  CPU: the multi-dimensional
oracle (`_milstein_nd` with the Stratonovich integrals) on the reference's SABR
test and on the zero-gradient identity the device route rests on."""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import milstein as omilstein



def _tff():
  import tff_b200 as tff
  return tff


@pytest.mark.gpu
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


@pytest.mark.gpu
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


@pytest.mark.gpu
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


@pytest.mark.gpu
def test_milstein_argument_errors():
  tff = _tff()
  from tff_b200.models import closures
  drift, vol = closures.gbm_closures(0.1, 0.2)
  mv = tff.models.MultivariateGeometricBrownianMotion(
      2, means=np.array([0.1, 0.2]), volatilities=np.array([0.2, 0.3]),
      corr_matrix=np.array([[1.0, 0.5], [0.5, 1.0]]), dtype=np.float64)
  with pytest.raises(NotImplementedError):               # state-dependent volatility matrix
    tff.models.milstein_sampling.sample(dim=2, drift_fn=mv.drift_fn(), volatility_fn=mv.volatility_fn(),
                                        times=[1.0], time_step=0.1, initial_state=np.ones(2))
  with pytest.raises(ValueError):
    tff.models.milstein_sampling.sample(dim=1, drift_fn=drift, volatility_fn=vol, times=[1.0])
  with pytest.raises(ValueError):
    tff.models.milstein_sampling.sample(dim=1, drift_fn=drift, volatility_fn=vol, times=[1.0],
                                        time_step=0.1, num_time_steps=10)


# ------------------------------------------------ multi-dimensional scheme ----
def _sabr(beta=0.5, volvol=1.0, rho=0.2):
  """milstein_sampling_test.py:233-277: dF = v F^beta dW_F, dv = volvol v dW_v, corr rho."""
  def vol_fn(t, x):
    del t
    f, v = x[..., 0], x[..., 1]
    fb = np.power(np.maximum(f, 0.0), beta)
    m = np.zeros(x.shape + (2,), dtype=x.dtype)
    m[..., 0, 0] = v * fb * np.sqrt(1 - rho**2)
    m[..., 0, 1] = v * fb * rho
    m[..., 1, 1] = volvol * v
    m[f <= 0.0] = 0.0
    return m

  def grad_fn(t, x):
    del t
    f, v = x[..., 0], x[..., 1]
    ok = f > 0.0
    fs = np.where(ok, f, 1.0)
    fb = np.power(fs, beta)
    dfb = beta * np.power(fs, beta - 1)
    g0 = np.zeros(x.shape + (2,), dtype=x.dtype)      # d vol / d f
    g0[..., 0, 0] = v * dfb * np.sqrt(1 - rho**2)
    g0[..., 0, 1] = v * dfb * rho
    g1 = np.zeros(x.shape + (2,), dtype=x.dtype)      # d vol / d v
    g1[..., 0, 0] = fb * np.sqrt(1 - rho**2)
    g1[..., 0, 1] = fb * rho
    g1[..., 1, 1] = volvol
    g0[~ok] = 0.0
    g1[~ok] = 0.0
    return [g0, g1]
  return vol_fn, grad_fn


def test_oracle_nd_sabr_statistics_match_euler():
  # milstein_sampling_test.py:279-323: mean / std of all Milstein paths against the Euler paths
  from oracle import euler as oeuler
  vol_fn, grad_fn = _sabr()
  times = np.array([0.0, 0.1, 0.21, 0.32, 0.43, 0.55])
  x0 = np.array([0.1, 0.2])
  paths = omilstein.sample(
      dim=2, drift_fn=lambda t, x: np.zeros_like(x), volatility_fn=vol_fn, grad_volatility_fn=grad_fn,
      times=times, num_samples=1000, initial_state=x0,
      random_type=odraws.RandomType.STATELESS_ANTITHETIC, time_step=0.01, seed=[1, 42])
  assert paths.shape == (1000, 6, 2) and np.isfinite(paths).all()
  euler = oeuler.sample(2, lambda t, x: np.zeros_like(x), vol_fn, times, time_step=0.01,
                        num_samples=10000, initial_state=x0,
                        random_type=odraws.RandomType.STATELESS_ANTITHETIC, seed=[1, 42],
                        dtype=np.float64)
  np.testing.assert_allclose((paths.mean(), paths.std()), (euler.mean(), euler.std()),
                             rtol=0.05, atol=0.05)


def _affine_nd(dim, dtype):
  rng = np.random.default_rng(7)
  a1 = (-0.5 * np.eye(dim) + 0.1 * rng.standard_normal((dim, dim))).astype(dtype)
  a0 = rng.standard_normal(dim).astype(dtype) * 0.1
  b = (0.2 * np.eye(dim) + 0.05 * rng.standard_normal((dim, dim))).astype(dtype)
  return a0, a1, b


@pytest.mark.parametrize('dim', [2, 3])
def test_oracle_nd_with_state_independent_volatility_is_the_euler_step(dim):
  # the identity the device route rests on: zero volatility gradient -> x + dt a + B dW on the
  # first `dim` columns of the Milstein draw tensor (dim + 3 dim order normals per step)
  dtype = np.float64
  a0, a1, b = _affine_nd(dim, dtype)
  times, n, order = np.array([0.3, 0.7]), 257, 4
  kw = dict(times=times, num_samples=n, initial_state=np.full(dim, 0.5), time_step=0.1,
            random_type=odraws.RandomType.STATELESS, seed=[3, 9])
  got = omilstein.sample(
      dim=dim, drift_fn=lambda t, x: a0 * (1 + t) + x @ a1.T,
      volatility_fn=lambda t, x: np.broadcast_to(b * (1 + t), x.shape + (dim,)),
      grad_volatility_fn=lambda t, x: [np.zeros(x.shape + (dim,)) for _ in range(dim)],
      stratonovich_order=order, **kw)
  from oracle import grid as ogrid
  all_times, keep, _ = ogrid.prepare_grid(times=times, time_step=np.float64(0.1), dtype=dtype)
  draws = odraws.generate_mc_normal_draws(
      num_normal_draws=dim + 3 * dim * order, num_time_steps=all_times.shape[0] - 1,
      num_sample_paths=n, random_type=odraws.RandomType.STATELESS, dtype=dtype, seed=[3, 9])
  x = np.full((n, dim), 0.5)
  out = []
  for i in range(all_times.shape[0] - 1):
    t, dt = all_times[i + 1], all_times[i + 1] - all_times[i]
    x = x + dt * (a0 * (1 + t) + x @ a1.T) + (draws[i][:, :dim] * np.sqrt(dt)) @ (b * (1 + t)).T
    if keep[i + 1]:
      out.append(x)
  np.testing.assert_allclose(got, np.stack(out, 1), rtol=1e-13, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize('rt', ['STATELESS', 'SOBOL', 'STATELESS_ANTITHETIC'])
@pytest.mark.parametrize('dim', [2, 3])
def test_sample_paths_nd_state_independent_volatility(dim, rt):
  tff = _tff()
  import torch
  dtype = np.float64
  a0, a1, b = _affine_nd(dim, dtype)
  ta0, ta1, tb = (torch.as_tensor(v) for v in (a0, a1, b))   # the callables are probed on the host
  times = np.array([0.0, 0.25, 0.6, 1.0])
  n = 2048
  seed = None if rt == 'SOBOL' else [5, 11]
  kw = dict(dim=dim, times=times, num_samples=n, initial_state=np.full(dim, 0.5), time_step=0.05,
            seed=seed, stratonovich_order=3)
  got = tff.models.milstein_sampling.sample(
      drift_fn=lambda t, x: ta0 * (1 + t) + x @ ta1.T,
      volatility_fn=lambda t, x: (tb * (1 + t)).expand(x.shape + (dim,)),
      random_type=tff.math.random.RandomType[rt], **kw).cpu().numpy()
  want = omilstein.sample(
      drift_fn=lambda t, x: a0 * (1 + t) + x @ a1.T,
      volatility_fn=lambda t, x: np.broadcast_to(b * (1 + t), x.shape + (dim,)),
      grad_volatility_fn=lambda t, x: [np.zeros(x.shape + (dim,)) for _ in range(dim)],
      random_type=odraws.RandomType[rt], **kw)
  assert got.shape == (n, 4, dim)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
