"""`mv_normal_sample` with `covariance_matrix` / `scale_matrix`, and `uniform` (CPU).

1. The reference's own tests (`math/random_ops/multivariate_normal_test.py:31-281`)
   run on the oracle (`oracle/draws.py`), same sample counts / seeds / tolerances.
2. The mirror's host wiring (`tff_b200.math.random.multivariate_normal`: batch
   shapes, the quasi-random transpose / reshape / permute, `scale · z + mean`,
   antithetic reflection about the mean) against the oracle, with the device
   generators replaced by CPU stand-ins built from the oracle's streams -- the
   kernels behind them are compared with the oracle in `tests/test_gpu_parity.py`
   and `tests/test_halton.py`.
"""
import numpy as np
import pytest
import torch

from oracle import draws as odraws
from oracle import halton as ohalton
from oracle import philox as ophilox
from oracle import sobol as osobol

RT = odraws.RandomType
MEAN = np.array([[1.0, 0.1], [0.1, 1.0]])
SCALE = np.array([[0.4, -0.1], [0.22, 1.38]])
COVAR2 = np.array([[[0.9, -0.1], [-0.1, 1.0]], [[1.1, -0.3], [-0.3, 0.6]]])


def _cov(x):
  return np.cov(x, rowvar=False)


# ------------------------------------------------ reference tests on the oracle
def test_shapes():
  # multivariate_normal_test.py:31-39
  assert odraws.mv_normal_sample([2, 4], mean=[0.2, 0.1], seed=1).shape == (2, 4, 2)
  assert odraws.mv_normal_sample([2, 4], mean=[[0.2, 0.1], [0., -0.1], [0., 0.1]], seed=1).shape == (2, 4, 3, 2)


def test_mean_default():
  # :41-53
  covar = np.array([[1.0, 0.1], [0.1, 1.0]])
  sample = odraws.mv_normal_sample([40000], covariance_matrix=covar, seed=1234)
  assert sample.shape == (40000, 2)
  np.testing.assert_allclose(sample.mean(axis=0), [0.0, 0.0], atol=1e-2)
  np.testing.assert_allclose(_cov(sample), covar, atol=2e-2)


def test_covariance_default():
  # :55-67
  sample = odraws.mv_normal_sample([10000], mean=MEAN, seed=4)
  assert sample.shape == (10000, 2, 2)
  np.testing.assert_array_almost_equal(sample.mean(axis=0), MEAN, decimal=1)
  for i in range(2):
    np.testing.assert_array_almost_equal(_cov(sample[:, i, :]), np.eye(2), decimal=1)


@pytest.mark.parametrize('random_type,seed', [(RT.PSEUDO, 4567), (RT.STATELESS, [1, 4567]), (RT.HALTON, None)])
def test_general_mean_covariance(random_type, seed):
  # :69-134 (PSEUDO, STATELESS; HALTON from `test_dynamic_shapes`)
  size = 30000
  sample = odraws.mv_normal_sample([size], mean=MEAN, covariance_matrix=COVAR2, random_type=random_type,
                                   seed=seed)
  assert sample.shape == (size, 2, 2)
  np.testing.assert_array_almost_equal(sample.mean(axis=0), MEAN, decimal=1)
  for i in range(2):
    np.testing.assert_array_almost_equal(_cov(sample[:, i, :]), COVAR2[i], decimal=1)


def test_mean_and_scale():
  # :136-153
  size = 30000
  sample = odraws.mv_normal_sample([size], mean=MEAN, scale_matrix=SCALE, seed=7534)
  assert sample.shape == (size, 2, 2)
  np.testing.assert_array_almost_equal(sample.mean(axis=0), MEAN, decimal=1)
  for i in range(2):
    np.testing.assert_array_almost_equal(_cov(sample[:, i, :]), SCALE @ SCALE.T, decimal=1)


@pytest.mark.parametrize('random_type', [RT.SOBOL, RT.HALTON_RANDOMIZED])
def test_mean_default_quasi(random_type):
  # :155-168 (SOBOL), :188-201 (HALTON_RANDOMIZED), skip = 1000
  covar = np.array([[1.0, 0.1], [0.1, 1.0]])
  sample = odraws.mv_normal_sample([10000], covariance_matrix=covar, random_type=random_type, skip=1000,
                                   seed=None if random_type == RT.SOBOL else 3)
  assert sample.shape == (10000, 2)
  np.testing.assert_allclose(sample.mean(axis=0), [0.0, 0.0], atol=1e-2)
  np.testing.assert_allclose(_cov(sample), covar, atol=2e-2)


@pytest.mark.parametrize('random_type,row', [(RT.SOBOL, 1), (RT.HALTON, 2)])
def test_mean_and_scale_quasi(random_type, row):
  # :170-186 (SOBOL), :221-238 (HALTON)
  mean = np.array([[1.0, 0.1], [0.1, 1.0], [2.0, 0.3], [0., 0.]])
  sample_shape = [2, 3, 5000]
  sample = odraws.mv_normal_sample(sample_shape, mean=mean, scale_matrix=SCALE, random_type=random_type)
  assert sample.shape == tuple(sample_shape) + (4, 2)
  np.testing.assert_array_almost_equal(sample.mean(axis=(0, 1, 2)), mean, decimal=1)
  for i in range(4):
    np.testing.assert_array_almost_equal(_cov(sample[0, row, :, i, :]), SCALE @ SCALE.T, decimal=1)


@pytest.mark.parametrize('random_type,seed', [(RT.PSEUDO_ANTITHETIC, 42), (RT.STATELESS_ANTITHETIC, [1, 42])])
def test_mean_and_scale_antithetic(random_type, seed):
  # :249-281
  size = 30000
  sample = odraws.mv_normal_sample([size], mean=MEAN, scale_matrix=SCALE, random_type=random_type, seed=seed)
  assert sample.shape == (size, 2, 2)
  half = size // 2
  np.testing.assert_allclose((sample[:half] + sample[half:]) / 2, MEAN + np.zeros([half, 2, 2]), 1e-10, 1e-10)
  np.testing.assert_array_almost_equal(sample[:half].mean(axis=0), MEAN, decimal=1)
  for i in range(2):
    np.testing.assert_array_almost_equal(_cov(sample[:half, i, :]), SCALE @ SCALE.T, decimal=1)


def test_antithetic_sample_requires_even_dim():
  # :283-293
  with pytest.raises(ValueError):
    odraws.mv_normal_sample([11, 100], mean=MEAN, scale_matrix=SCALE, random_type=RT.PSEUDO_ANTITHETIC)


# --------------------------------------------------- the mirror's host wiring
@pytest.fixture
def host_generators(monkeypatch):
  """Replace the device fills of `tff_b200.math.random` by CPU tensors holding the
  oracle's streams, so that only the host wiring of `multivariate_normal` runs."""
  from tff_b200 import _tensor
  from tff_b200.math.random import halton, philox, sobol
  monkeypatch.setattr(_tensor, 'device', lambda: torch.device('cpu'))
  monkeypatch.setattr(philox, 'normal', lambda shape, dtype=None, seed=None: torch.from_numpy(
      ophilox.stateful_normal(tuple(shape), seed, np.dtype(dtype))))
  monkeypatch.setattr(philox, 'stateless_normal', lambda shape, seed, dtype=None: torch.from_numpy(
      ophilox.stateless_normal(tuple(shape), seed, np.dtype(dtype))))
  monkeypatch.setattr(sobol, 'sample_normal', lambda dim, n, skip=0, dtype=None: torch.from_numpy(
      odraws._erfinv_times_sqrt2(osobol.sample(dim, n, skip=skip, dtype=dtype), dtype)))

  def halton_normal(dim, n, skip=0, dtype=None, randomized=False, seed=None, randomization_params=None):
    assert randomization_params is None
    u = ohalton.sample(dim, sequence_indices=np.arange(skip, skip + n), dtype=dtype, randomized=randomized,
                       seed=seed)
    return torch.from_numpy(odraws._erfinv_times_sqrt2(u, dtype))
  monkeypatch.setattr(halton, 'sample_normal', halton_normal)


_CASES = [
    dict(sample_shape=[64], mean=MEAN, scale_matrix=SCALE),
    dict(sample_shape=[64], mean=MEAN, covariance_matrix=COVAR2),
    dict(sample_shape=[64], covariance_matrix=COVAR2[0]),
    dict(sample_shape=[64], scale_matrix=SCALE),
    dict(sample_shape=[2, 3, 16], mean=np.array([[1.0, 0.1], [0.1, 1.0], [2.0, 0.3], [0., 0.]]),
         scale_matrix=SCALE),
    dict(sample_shape=[4, 8], mean=np.array([0.2, 0.1, -0.4])),
]


@pytest.mark.parametrize('random_type,seed', [
    ('PSEUDO', 11), ('STATELESS', [1, 4567]), ('PSEUDO_ANTITHETIC', 42), ('STATELESS_ANTITHETIC', [1, 42]),
    ('SOBOL', None), ('HALTON', None), ('HALTON_RANDOMIZED', 7889)])
@pytest.mark.parametrize('case', range(len(_CASES)))
def test_host_wiring_of_multivariate_normal_equals_the_oracle(host_generators, case, random_type, seed):
  import tff_b200 as tff
  kw = dict(_CASES[case])
  sample_shape = kw.pop('sample_shape')
  extra = {'skip': 5} if random_type in ('SOBOL', 'HALTON', 'HALTON_RANDOMIZED') else {}
  got = tff.math.random.mv_normal_sample(sample_shape, random_type=getattr(tff.math.random.RandomType, random_type),
                                         seed=seed, **kw, **extra)
  want = odraws.mv_normal_sample(sample_shape, random_type=getattr(RT, random_type), seed=seed, **kw, **extra)
  assert tuple(got.shape) == want.shape and got.dtype == torch.float64
  np.testing.assert_allclose(got.numpy(), want, rtol=1e-13, atol=1e-13)


@pytest.fixture
def host_uniform_generators(monkeypatch):
  from tff_b200.math.random import halton, philox, sobol
  monkeypatch.setattr(philox, 'uniform', lambda shape, dtype=None, seed=None: torch.from_numpy(
      ophilox.stateful_uniform(list(shape), seed, np.dtype(dtype))))
  monkeypatch.setattr(philox, 'stateless_uniform', lambda shape, seed, dtype=None: torch.from_numpy(
      ophilox.stateless_uniform(list(shape), seed, np.dtype(dtype))))
  monkeypatch.setattr(sobol, 'sample', lambda dim, num_results, skip=0, dtype=None: torch.from_numpy(
      osobol.sample(dim, num_results, skip=skip, dtype=dtype)))

  def halton_sample(dim, num_results=None, sequence_indices=None, randomized=True, randomization_params=None,
                    seed=None, validate_args=False, dtype=None, name=None):
    assert randomization_params is None and num_results is None
    return torch.from_numpy(ohalton.sample(dim, sequence_indices=np.asarray(sequence_indices), dtype=dtype,
                                           randomized=randomized, seed=seed)), None
  monkeypatch.setattr(halton, 'sample', halton_sample)


@pytest.mark.parametrize('random_type,seed', [
    ('PSEUDO', 101), ('STATELESS', [2, 2]), ('SOBOL', None), ('HALTON', None), ('HALTON_RANDOMIZED', 7),
    ('STATELESS_ANTITHETIC', [1, 2])])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_host_wiring_of_uniform_equals_the_oracle(host_uniform_generators, random_type, seed, dtype):
  # math/random_ops/uniform.py:92-153; STATELESS_ANTITHETIC falls into the reference's
  # quasi-random `else` branch and yields the plain Halton sequence
  import tff_b200 as tff
  extra = {} if random_type in ('PSEUDO', 'STATELESS') else {'skip': 1000}
  got = tff.math.random.uniform(5, [2, 3, 10], random_type=getattr(tff.math.random.RandomType, random_type),
                                seed=seed, dtype=dtype, **extra)
  want = odraws.uniform(5, [2, 3, 10], random_type=getattr(RT, random_type), seed=seed, dtype=dtype, **extra)
  assert tuple(got.shape) == (2, 3, 10, 5) == want.shape and got.numpy().dtype == dtype
  np.testing.assert_array_equal(got.numpy(), want)
  if random_type == 'STATELESS_ANTITHETIC':
    np.testing.assert_array_equal(want, odraws.uniform(5, [2, 3, 10], random_type=RT.HALTON, dtype=dtype, skip=1000))


def test_uniform_errors():
  import tff_b200 as tff
  rt = tff.math.random.RandomType
  with pytest.raises(ValueError):
    tff.math.random.uniform(2, [4], random_type=rt.STATELESS)
  with pytest.raises(NotImplementedError):
    tff.math.random.uniform(2, [4], random_type=rt.PSEUDO_ANTITHETIC, seed=1)
  with pytest.raises(ValueError):
    odraws.uniform(2, [4], random_type=RT.STATELESS)
  with pytest.raises(NotImplementedError):
    odraws.uniform(2, [4], random_type=RT.PSEUDO_ANTITHETIC, seed=1)
