"""Halton sequence, plain and Owen-randomized (SURVEY 8f-4;
`math/random_ops/halton/halton_impl.py:59-379`, `halton_test.py:30-300`).

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


# ------------------------------------------------------------- randomized ----
def test_oracle_randomized_structure():
  """What the reference's randomized tests assert that is not statistical
  (`halton_test.py:205-300`): determinism in the seed, reuse of the returned
  params, access by index; plus the structure of `_get_permutations` / the zero
  correction."""
  dim, n = 7, 400
  x, (perms, zc) = ohalton.sample(dim, num_results=n, dtype=np.float64, randomized=True,
                                  seed=1925, return_params=True)
  assert x.shape == (n, dim) and x.min() >= 0.0 and x.max() < 1.0
  np.testing.assert_array_equal(
      x, ohalton.sample(dim, num_results=n, dtype=np.float64, randomized=True, seed=1925))
  assert not np.array_equal(
      x, ohalton.sample(dim, num_results=n, dtype=np.float64, randomized=True, seed=1926))
  # halton_test.py:274-287: a second seed is ignored when params are supplied
  np.testing.assert_array_equal(
      x, ohalton.sample(dim, num_results=n, dtype=np.float64, randomized=True, seed=62278,
                        randomization_params=(perms, zc)))
  # halton_test.py:225-240: batches by sequence_indices equal the full sample
  np.testing.assert_array_equal(
      x[100:200], ohalton.sample(dim, sequence_indices=np.arange(100, 200), dtype=np.float64,
                                 randomized=True, seed=1925))
  # every digit position of every axis holds a permutation of range(p)
  pr = ohalton.primes(dim)
  table = perms.reshape(ohalton.num_coeffs(np.float64), int(pr.sum()))
  off = 0
  for p in pr:
    block = table[:, off:off + p]
    np.testing.assert_array_equal(np.sort(block, axis=1), np.tile(np.arange(p), (table.shape[0], 1)))
    off += p
  sizes = ohalton.max_sizes_by_axes(dim, np.float64).reshape(-1)
  assert np.all(zc >= 0) and np.all(zc < pr.astype(np.float64)**-sizes)
  # Owen scrambling keeps the radical-inverse structure: the first p^k points of axis d
  # fall one in each interval of width p^-k
  for d, p in enumerate(pr[:4]):
    k = 2
    cells = np.floor(x[:p**k, d] * p**k).astype(int)
    assert sorted(cells.tolist()) == list(range(p**k))
  # halton_test.py:139-176 in spirit: the randomized estimate of an integral is unbiased
  est = [ohalton.sample(3, num_results=500, dtype=np.float64, randomized=True,
                        seed=121117 + i).prod(axis=1).mean() for i in range(20)]
  assert abs(np.mean(est) - 0.125) < 3e-4


def test_library_permutations_match_oracle():
  """tqf_halton_permutations is host code: checked without a GPU, bit for bit."""
  import ctypes as C
  from tff_b200 import _lib
  for dim, seed, dtype in ((5, 1925, np.float64), (12, 7, np.float32), (1, 0, np.float64)):
    pr = ohalton.primes(dim)
    nc = ohalton.num_coeffs(dtype)
    got = np.empty(nc * int(pr.sum()), dtype=np.int32)
    _lib.check(_lib.lib().tqf_halton_permutations(seed, pr.ctypes.data, dim, nc, got.ctypes.data))
    np.testing.assert_array_equal(got, ohalton.get_permutations(nc, pr, seed).reshape(-1))


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_gpu_randomized_halton_matches_oracle(dtype):
  import tff_b200 as tff
  halton = tff.math.random.halton
  tol = dict(rtol=4e-16, atol=2e-16) if dtype == np.float64 else dict(rtol=2.5e-7, atol=1.2e-7)
  for dim, start, n, seed in ((2, 0, 50, 11), (40, 12345, 3000, 1729), (300, 7, 33, 127)):
    idx = np.arange(start, start + n)
    got, params = halton.sample(dim, sequence_indices=idx, seed=seed, dtype=dtype)   # randomized
    want, (perms, zc) = ohalton.sample(dim, sequence_indices=idx, dtype=dtype, randomized=True,
                                       seed=seed, return_params=True)
    assert _np(got).dtype == dtype and tuple(got.shape) == (n, dim)
    np.testing.assert_allclose(_np(got), want, **tol)
    np.testing.assert_array_equal(_np(params.perms), perms)
    np.testing.assert_allclose(_np(params.zero_correction), zc, rtol=1e-6 if dtype == np.float32 else 1e-15)
    # params reuse: the seed is ignored (halton_test.py:274-287)
    again, _ = halton.sample(dim, sequence_indices=idx, seed=seed + 5, dtype=dtype,
                             randomization_params=params)
    np.testing.assert_array_equal(_np(again), _np(got))
  # unseeded: random, valid, and reproducible through its params
  a, pa = halton.sample(4, num_results=64, dtype=dtype)
  b, _ = halton.sample(4, num_results=64, dtype=dtype, randomization_params=pa)
  np.testing.assert_array_equal(_np(a), _np(b))
  assert _np(a).min() >= 0 and _np(a).max() < 1


@pytest.mark.gpu
def test_gpu_halton_randomized_random_type():
  import tff_b200 as tff
  from tff_b200.models import closures
  rt = tff.math.random.RandomType.HALTON_RANDOMIZED
  got = _np(tff.math.random.uniform(5, [100], random_type=rt, skip=1000, seed=42, dtype=np.float64))
  want = ohalton.sample(5, sequence_indices=np.arange(1000, 1100), dtype=np.float64,
                        randomized=True, seed=42)
  np.testing.assert_allclose(got, want, rtol=4e-16, atol=2e-16)
  mean = np.zeros(6)
  got = _np(tff.math.random.mv_normal_sample([500], mean=mean, random_type=rt, skip=3, seed=9))
  want = odraws.mv_normal_sample([500], mean, random_type=odraws.RandomType.HALTON_RANDOMIZED,
                                 skip=3, seed=9)
  np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)
  mu, sigma = 0.03, 0.2
  drift, vol = closures.gbm_closures(mu, sigma)
  kw = dict(num_samples=2000, initial_state=np.array([1.5]), time_step=0.1, skip=5, seed=77,
            dtype=np.float64)
  got = _np(tff.models.euler_sampling.sample(1, drift, vol, [0.5, 1.0], random_type=rt, **kw))
  want = oeuler.sample(1, lambda t, x: mu * x, lambda t, x: (sigma * x)[..., None], [0.5, 1.0],
                       random_type=odraws.RandomType.HALTON_RANDOMIZED, **kw)
  np.testing.assert_allclose(got, want, rtol=1e-11)
  # the randomized sequence integrates: E[S_T] = 1.5 exp(mu)
  assert abs(got[:, 1, 0].mean() - 1.5 * np.exp(mu)) < 5e-3
