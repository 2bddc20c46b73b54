"""Non-randomized Halton sequence (SURVEY 8f-4;
`math/random_ops/halton/halton_impl.py:59-288`, `halton_test.py:30-75`).

CPU: the oracle against the reference's known values and against exact rational
radical inverses.  GPU: the fill kernel against the oracle, `uniform` /
`mv_normal_sample` / `euler_sampling.sample` with `RandomType.HALTON`."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import halton as ohalton


def _radical_inverse(i, b):
  f, r = Fraction(1), Fraction(0)
  while i > 0:
    f /= b
    r += f * (i % b)
    i //= b
  return r


def test_oracle_known_values_small_bases():
  # halton_test.py:30-37 (and 39-48)
  expected = np.array([[1. / 2, 1. / 3], [1. / 4, 2. / 3], [3. / 4, 1. / 9],
                       [1. / 8, 4. / 9], [5. / 8, 7. / 9]], dtype=np.float32)
  np.testing.assert_allclose(ohalton.sample(2, num_results=5), expected, rtol=1e-6)
  # halton_test.py:50-61: access by index
  np.testing.assert_allclose(ohalton.sample(5, num_results=10),
                             ohalton.sample(5, sequence_indices=np.arange(10)), rtol=1e-6)
  assert ohalton.sample(3, num_results=10, dtype=np.float32).dtype == np.float32
  assert ohalton.sample(3, num_results=10, dtype=np.float64).dtype == np.float64
  with pytest.raises(ValueError):
    ohalton.sample(2)
  # the first 1000 primes end at 7919; digits per axis as _NUM_COEFFS_BY_DTYPE (24 / 54)
  assert ohalton.primes(1000)[-1] == 7919
  assert int(ohalton.max_sizes_by_axes(1, np.float32)[0, 0]) == 24
  assert int(ohalton.max_sizes_by_axes(1, np.float64)[0, 0]) == 54


@pytest.mark.parametrize('dtype,tol', [(np.float64, 4e-16), (np.float32, 2e-7)])
def test_oracle_equals_exact_radical_inverse(dtype, tol):
  dim, start, n = 40, 12345, 64
  got = ohalton.sample(dim, sequence_indices=np.arange(start, start + n), dtype=dtype)
  pr = ohalton.primes(dim)
  want = np.array([[float(_radical_inverse(i + 1, int(b))) for b in pr]
                   for i in range(start, start + n)])
  np.testing.assert_allclose(got, want, rtol=tol, atol=0)


def _np(t):
  return t.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_gpu_halton_matches_oracle(dtype):
  import tff_b200 as tff
  for dim, start, n in ((2, 0, 5), (40, 12345, 3000), (1000, 7, 33)):
    got, params = tff.math.random.halton.sample(dim, sequence_indices=np.arange(start, start + n),
                                                randomized=False, dtype=dtype)
    want = ohalton.sample(dim, sequence_indices=np.arange(start, start + n), dtype=dtype)
    assert params is None and _np(got).dtype == dtype and tuple(got.shape) == (n, dim)
    # same operations in the same order: equal up to the division's last bit
    np.testing.assert_allclose(_np(got), want, rtol=4e-16 if dtype == np.float64 else 2.5e-7)
  got, _ = tff.math.random.halton.sample(3, num_results=10, randomized=False, dtype=dtype)
  np.testing.assert_allclose(_np(got), ohalton.sample(3, num_results=10, dtype=dtype), rtol=1e-6)
  with pytest.raises(NotImplementedError):
    tff.math.random.halton.sample(3, num_results=10)            # randomized=True is the default
  with pytest.raises(ValueError):
    tff.math.random.halton.sample(3, randomized=False)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_gpu_uniform_and_normal_halton(dtype):
  import tff_b200 as tff
  rt = tff.math.random.RandomType.HALTON
  got = _np(tff.math.random.uniform(5, [100], random_type=rt, skip=1000, dtype=dtype))
  want = ohalton.sample(5, sequence_indices=np.arange(1000, 1100), dtype=dtype)
  np.testing.assert_allclose(got, want, rtol=4e-16 if dtype == np.float64 else 2.5e-7)
  mean = np.zeros(6, dtype=dtype)
  got = _np(tff.math.random.mv_normal_sample([500], mean=mean, random_type=rt, skip=3))
  want = odraws.mv_normal_sample([500], mean, random_type=odraws.RandomType.HALTON, skip=3)
  assert got.dtype == dtype and got.shape == (500, 6)
  np.testing.assert_allclose(got, want, rtol=1e-12 if dtype == np.float64 else 2e-5,
                             atol=1e-14 if dtype == np.float64 else 2e-6)


@pytest.mark.gpu
def test_gpu_euler_sample_with_halton_draws():
  import tff_b200 as tff
  from tff_b200.models import closures
  mu, sigma = 0.03, 0.2
  drift, vol = closures.gbm_closures(mu, sigma)
  kw = dict(num_samples=2000, initial_state=np.array([1.5]), time_step=0.1, skip=5,
            dtype=np.float64)
  got = _np(tff.models.euler_sampling.sample(1, drift, vol, [0.5, 1.0],
                                             random_type=tff.math.random.RandomType.HALTON, **kw))
  want = oeuler.sample(1, lambda t, x: mu * x, lambda t, x: (sigma * x)[..., None], [0.5, 1.0],
                       random_type=odraws.RandomType.HALTON, **kw)
  assert got.shape == want.shape == (2000, 2, 1)
  np.testing.assert_allclose(got, want, rtol=1e-12)
