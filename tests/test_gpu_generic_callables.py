"""The reference's own Euler-sampler tests (models/euler_sampling_test.py) run
against the engine through plain Python drift / volatility callables, which
the engine accepts when they are affine in the state (probed on the host)."""
import numpy as np
import pytest
import torch

from oracle import draws as odraws
from oracle import euler as oeuler

pytestmark = pytest.mark.gpu


def _tff():
  import tff_b200 as tff
  return tff


def _np(t):
  return t.detach().cpu().numpy()


@pytest.mark.parametrize('rt,seed', [('STATELESS', [1, 2]), ('SOBOL', None),
                                     ('PSEUDO_ANTITHETIC', 42)])
def test_sample_paths_1d(rt, seed):
  # euler_sampling_test.py:174-294: dX = mu sqrt(t) dt + (a t + b) dW
  tff = _tff()
  mu, a, b = 0.2, 0.4, 0.33

  def drift_fn(t, x):
    return mu * torch.sqrt(t) * torch.ones_like(x)

  def vol_fn(t, x):
    del x
    return (a * t + b) * torch.ones([1, 1], dtype=t.dtype)
  times = np.array([0.0, 0.1, 0.21, 0.32, 0.43, 0.55])
  n = 10000
  x0 = np.array([0.1])
  kw = dict(num_samples=n, initial_state=x0, seed=seed, time_step=0.01, dtype=np.float64)
  paths = _np(tff.models.euler_sampling.sample(
      1, drift_fn, vol_fn, times, random_type=tff.math.random.RandomType[rt], **kw))
  assert paths.shape == (n, 6, 1)
  means = paths.mean(axis=0)[:, 0]
  np.testing.assert_allclose(means, x0 + (2.0 / 3.0) * mu * times**1.5, rtol=1e-2, atol=1e-2)
  # times[0] == 0: the first recorded state is the initial state
  np.testing.assert_array_equal(paths[:, 0, 0], 0.1)
  want = oeuler.sample(1, lambda t, x: mu * np.sqrt(t) * np.ones_like(x),
                       lambda t, x: (a * t + b) * np.ones(x.shape + (1,)), times,
                       random_type=odraws.RandomType[rt], **kw)
  np.testing.assert_allclose(paths, want, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize('dim', [2, 3, 4])
def test_wiener_process_mean_and_covariance(dim):
  # euler_sampling_test.py:32-133: dX = dW, Cov(X_s, X_t) = min(s, t)
  tff = _tff()

  def drift_fn(_, x):
    return torch.zeros_like(x)

  def vol_fn(_, x):
    return torch.eye(dim, dtype=x.dtype).expand(x.shape[0], dim, dim)
  times = np.array([0.1, 0.2, 0.5])
  n = 20000
  paths = _np(tff.models.euler_sampling.sample(
      dim, drift_fn, vol_fn, times, num_samples=n, time_step=0.05,
      random_type=tff.math.random.RandomType.STATELESS, seed=[3, 4], dtype=np.float64))
  assert paths.shape == (n, 3, dim)
  np.testing.assert_allclose(paths.mean(axis=0), 0.0, atol=1e-2)
  for j in range(dim):
    cov = np.cov(paths[:, :, j], rowvar=False)
    np.testing.assert_allclose(cov, np.minimum.outer(times, times), rtol=5e-2, atol=1e-2)
  want = oeuler.sample(dim, lambda t, x: np.zeros_like(x),
                       lambda t, x: np.broadcast_to(np.eye(dim), x.shape + (dim,)), times,
                       num_samples=n, time_step=0.05, random_type=odraws.RandomType.STATELESS,
                       seed=[3, 4], dtype=np.float64)
  np.testing.assert_allclose(paths, want, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_sample_paths_2d(dtype):
  # euler_sampling_test.py:296-361: dX_i = mu_i sqrt(t) dt + S(t) dW with a
  # time-dependent (state-independent) matrix S, plus a linear drift term
  tff = _tff()
  mu = np.array([0.2, 0.7])
  a = np.array([[0.4, 0.1], [0.3, 0.2]])
  b = np.array([[0.33, -0.03], [0.21, 0.5]])
  k = np.array([[-0.5, 0.1], [0.0, -0.3]])

  def drift_fn(t, x):
    return torch.as_tensor(mu, dtype=x.dtype) * torch.sqrt(t) + x @ torch.as_tensor(k.T, dtype=x.dtype)

  def vol_fn(t, x):
    del x
    return torch.as_tensor(a, dtype=t.dtype) * t + torch.as_tensor(b, dtype=t.dtype)
  times = np.array([0.1, 0.21, 0.32, 0.43, 0.55])
  n = 4000
  x0 = np.array([0.1, -1.1])
  kw = dict(num_samples=n, initial_state=x0.astype(dtype), time_step=0.01, skip=100, dtype=dtype)
  paths = _np(tff.models.euler_sampling.sample(
      2, drift_fn, vol_fn, times, random_type=tff.math.random.RandomType.SOBOL, **kw))
  assert paths.shape == (n, 5, 2) and paths.dtype == dtype
  want = oeuler.sample(
      2, lambda t, x: (mu * np.sqrt(t) + x @ k.T).astype(x.dtype),
      lambda t, x: np.broadcast_to((a * t + b).astype(x.dtype), x.shape + (2,)), times,
      random_type=odraws.RandomType.SOBOL, **kw)
  if dtype == np.float64:
    np.testing.assert_allclose(paths, want, rtol=1e-11, atol=1e-13)
  else:
    np.testing.assert_allclose(paths, want, rtol=1e-4, atol=1e-5)


def test_state_dependent_volatility_1d_and_generic_ito_process():
  # generic_ito_process_test.py style: GenericItoProcess with lambdas, GBM form
  tff = _tff()
  mu, sigma = 0.05, 0.3
  process = tff.models.GenericItoProcess(
      1, lambda t, x: mu * x, lambda t, x: (sigma * x).unsqueeze(-1), dtype=np.float64)
  kw = dict(num_samples=3000, initial_state=np.array([2.0]), time_step=0.02, seed=[5, 6])
  paths = _np(process.sample_paths([0.5, 1.0], random_type=tff.math.random.RandomType.STATELESS, **kw))
  want = oeuler.sample(1, lambda t, x: mu * x, lambda t, x: (sigma * x)[..., None], [0.5, 1.0],
                       random_type=odraws.RandomType.STATELESS, dtype=np.float64, **kw)
  np.testing.assert_allclose(paths, want, rtol=1e-12)


def test_sample_paths_dtypes():
  # euler_sampling_test.py:484-501
  tff = _tff()
  for dtype in (np.float32, np.float64):
    paths = tff.models.euler_sampling.sample(
        dim=1, drift_fn=lambda t, x: torch.sqrt(t) * torch.ones_like(x),
        volatility_fn=lambda t, x: t * torch.ones([1, 1], dtype=t.dtype),
        times=[0.1, 0.2], num_samples=10, initial_state=[0.1], time_step=0.01, seed=123,
        dtype=dtype)
    assert _np(paths).dtype == dtype and tuple(paths.shape) == (10, 2, 1)


def test_argument_errors():
  # euler_sampling_test.py:503-570 and euler_sampling.py:254-263, 298-301
  tff = _tff()
  from tff_b200.models.euler_sampling import InvalidArgumentError
  drift = lambda _, x: torch.zeros_like(x)
  vol = lambda _, x: torch.ones_like(x).unsqueeze(-1)
  sample = tff.models.euler_sampling.sample
  with pytest.raises(InvalidArgumentError):
    sample(1, drift, vol, [0.1, 0.5, 2.0, 1.0], time_step=0.01, seed=42, validate_args=True,
           dtype=np.float64)
  with pytest.raises(InvalidArgumentError):
    sample(1, drift, vol, [0.1, 0.5, 1.0], times_grid=[0.1, 0.5, 1.0, 1.0], seed=42,
           validate_args=True, dtype=np.float64)
  draws = torch.zeros((100, 5, 1), dtype=torch.float64, device='cuda')
  with pytest.raises(InvalidArgumentError):
    sample(1, drift, vol, [0.1, 0.5, 1.0], normal_draws=draws, times_grid=[0.1, 0.5, 1.0],
           validate_args=True, dtype=np.float64)
  with pytest.raises(ValueError):
    sample(1, drift, vol, [1.0], dtype=np.float64)                      # no grid spec
  with pytest.raises(ValueError):
    sample(1, drift, vol, [1.0], time_step=0.1, num_time_steps=3, dtype=np.float64)
  with pytest.raises(ValueError):
    sample(2, drift, vol, [1.0], normal_draws=draws, time_step=0.2, dtype=np.float64)
  with pytest.raises(ValueError):                                        # odd antithetic count
    sample(1, drift, vol, [1.0], num_samples=11, time_step=0.5, seed=1,
           random_type=tff.math.random.RandomType.PSEUDO_ANTITHETIC, dtype=np.float64)
  with pytest.raises(NotImplementedError):                               # not affine
    sample(1, lambda t, x: torch.sin(x), vol, [1.0], time_step=0.5, seed=1, dtype=np.float64)
  # `watch_params` only steers how TensorFlow differentiates the loop: same forward paths
  rt = tff.math.random.RandomType.STATELESS
  a = sample(1, drift, vol, [1.0], time_step=0.5, seed=[1, 2], random_type=rt, num_samples=64,
             watch_params=[1.0], dtype=np.float64)
  b = sample(1, drift, vol, [1.0], time_step=0.5, seed=[1, 2], random_type=rt, num_samples=64,
             dtype=np.float64)
  assert torch.equal(a, b)


def test_repeated_price_calls_are_bound_by_content():
  # euler_sampling.price recognises a repeated call by the CONTENT of its arguments and runs it
  # as one FFI call (tqf_plan_price_host); any changed number must give the new answer
  tff = _tff()
  from tff_b200 import engine
  from tff_b200.models import closures
  from tff_b200.models import euler_sampling as es
  r, sigma = 0.03, 0.1
  d, v = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
  process = tff.models.GenericItoProcess(1, d, v, dtype=np.float64)
  x0 = np.array([np.log(700.0)])
  rt = tff.math.random.RandomType.PSEUDO_ANTITHETIC

  def call(strike, seed=42, n=20000, stats=False):
    return process.price([1.0], [engine.european_call(strike, log_state=True, scale=np.exp(-r))],
                         num_samples=n, initial_state=x0, random_type=rt, seed=seed, time_step=0.01,
                         return_stats=stats)
  es._CALLS.clear()
  first = call(650.0)
  assert len(es._CALLS) == 1
  again = call(650.0)
  np.testing.assert_array_equal(first, again)
  other = call(680.0)
  assert len(es._CALLS) == 2 and other[0] < first[0]
  np.testing.assert_array_equal(call(650.0), first)
  assert call(650.0, seed=43)[0] != first[0]
  mean, stderr, bad = call(650.0, stats=True)
  kept = bad.copy()
  call(680.0, stats=True)
  np.testing.assert_array_equal(bad, kept)                 # results do not alias the bound buffers
  np.testing.assert_array_equal(mean, first)
  # a destroyed plan is rebuilt, not reused
  engine.clear_plan_cache()
  np.testing.assert_array_equal(call(650.0), first)
  # callables of time are never bound
  d2, v2 = closures.affine_closures(lambda t: 0.02 + 0 * t, 0.0, sigma)
  p2 = tff.models.GenericItoProcess(1, d2, v2, dtype=np.float64)
  before = len(es._CALLS)
  p2.price([1.0], [engine.european_call(650.0, log_state=True)], num_samples=1000, initial_state=x0,
           random_type=rt, seed=1, time_step=0.1)
  assert len(es._CALLS) == before


def test_normal_draws_from_a_foreign_dlpack_exporter():
  # every tensor argument accepts any object that speaks the DLPack protocol (a tf.Tensor through
  # tf.experimental.dlpack, a cupy array, ...): an exporter that is not a torch.Tensor goes through
  # `__dlpack__` / `__dlpack_device__` and is consumed in place (no copy through the host)
  tff = _tff()
  from tff_b200.models import closures

  class Foreign:
    def __init__(self, t):
      self._t = t
      self.shape = tuple(t.shape)

    def __dlpack__(self, stream=None):
      return self._t.__dlpack__(stream=stream) if stream is not None else self._t.__dlpack__()

    def __dlpack_device__(self):
      return self._t.__dlpack_device__()

  d, v = closures.affine_closures(0.02, -0.3, 0.15, 0.25)
  draws = torch.randn((512, 10, 1), dtype=torch.float64, device='cuda',
                      generator=torch.Generator(device='cuda').manual_seed(5))
  kw = dict(times=[0.5, 1.0], times_grid=np.linspace(0.0, 1.0, 11), initial_state=np.array([1.3]),
            dtype=np.float64)
  a = tff.models.euler_sampling.sample(1, d, v, normal_draws=draws, **kw)
  b = tff.models.euler_sampling.sample(1, d, v, normal_draws=Foreign(draws), **kw)
  assert torch.equal(a, b) and a.shape == (512, 2, 1)
