"""C4 on the 5th-generation tensor cores: `mvgbm_tc5_kernel` (tcgen05.mma.kind::tf32,
accumulator and A operand in tensor memory) against the oracle and against the
mma.sync kernel it is A/B-tested with (`TQF_MVGBM_TC5` selects per launch)."""
import os

import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import models as omodels

pytestmark = pytest.mark.gpu


def _setup(dim, dtype):
  import tff_b200 as tff
  means = np.full(dim, 0.03, dtype)
  vols = np.linspace(0.1, 0.4, dim).astype(dtype)
  corr = (0.3 + 0.7 * np.eye(dim)).astype(dtype)
  model = tff.models.MultivariateGeometricBrownianMotion(dim, means=means, volatilities=vols,
                                                         corr_matrix=corr, dtype=dtype)
  closures = omodels.mvgbm_closures(means, vols, corr, dtype)
  return tff, model, closures, (100.0 * np.ones(dim)).astype(dtype)


class _Tc5:
  def __init__(self, on):
    self.on = on

  def __enter__(self):
    self.old = os.environ.get('TQF_MVGBM_TC5')
    os.environ['TQF_MVGBM_TC5'] = '1' if self.on else '0'

  def __exit__(self, *a):
    if self.old is None:
      del os.environ['TQF_MVGBM_TC5']
    else:
      os.environ['TQF_MVGBM_TC5'] = self.old


@pytest.mark.parametrize('n,steps,skip', [(4096, 12, 0), (1000, 5, 0), (777, 3, 12345), (128, 1, 0),
                                          (1, 2, 7)])
def test_tc5_basket_price_matches_oracle(n, steps, skip):
  from tff_b200 import engine
  dim, dtype = 64, np.float32
  tff, model, (odrift, ovol), x0 = _setup(dim, dtype)
  kw = dict(num_samples=n, initial_state=x0, num_time_steps=steps)
  payoffs = [engine.european_call(100.0, component=-1), engine.european_put(100.0, component=-1),
             engine.identity(component=5), engine.identity(component=63), engine.identity(component=0)]
  with _Tc5(True):
    got = model.price_euler([1.0], payoffs, random_type=tff.math.random.RandomType.SOBOL, skip=skip, **kw)
  with _Tc5(False):
    legacy = model.price_euler([1.0], payoffs, random_type=tff.math.random.RandomType.SOBOL, skip=skip, **kw)
  paths = oeuler.sample(dim, odrift, ovol, [1.0], random_type=odraws.RandomType.SOBOL, dtype=dtype,
                        skip=skip, **kw)[:, 0, :].astype(np.float64)
  basket = paths.mean(axis=1)
  want = [np.maximum(basket - 100, 0).mean(), np.maximum(100 - basket, 0).mean(), paths[:, 5].mean(),
          paths[:, 63].mean(), paths[:, 0].mean()]
  np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-4)
  np.testing.assert_allclose(got, legacy, rtol=2e-5, atol=1e-4)


def test_tc5_c4_shape_252_steps_has_no_systematic_error():
  # the 252-step bias test of test_gpu_mvgbm.py through identity payoffs: per-asset means of
  # 512 paths within 2e-6 of the oracle's (a rounding error common to all paths and steps shows
  # up here and not in a short test)
  from tff_b200 import engine
  dim, dtype, n = 64, np.float32, 512
  tff, model, (odrift, ovol), x0 = _setup(dim, dtype)
  kw = dict(num_samples=n, initial_state=x0, num_time_steps=252)
  want = oeuler.sample(dim, odrift, ovol, [1.0], random_type=odraws.RandomType.SOBOL,
                       dtype=dtype, **kw)[:, 0, :].astype(np.float64).mean(axis=0)
  got = np.zeros(dim)
  with _Tc5(True):
    for lo in range(0, dim, 8):
      pay = [engine.identity(component=i) for i in range(lo, lo + 8)]
      got[lo:lo + 8] = model.price_euler([1.0], pay, random_type=tff.math.random.RandomType.SOBOL, **kw)
  np.testing.assert_allclose(got, want, rtol=2e-6)


def test_tc5_counts_non_finite_paths_like_the_mma_kernel():
  # float32 Sobol u == 1.0 at index 18 684 944 in dimension 12 (step 0, asset 12)
  from tff_b200 import engine
  tff, model, _, x0 = _setup(64, np.float32)
  out = {}
  for on in (True, False):
    with _Tc5(on):
      out[on] = model.price_euler(
          [1.0], [engine.identity(component=12), engine.european_call(100.0, component=-1)],
          initial_state=x0, num_samples=256, random_type=tff.math.random.RandomType.SOBOL,
          skip=18684944 - 100, num_time_steps=1, return_stats=True)
  assert out[True][2][0] >= 1
  np.testing.assert_array_equal(out[True][2], out[False][2])
  np.testing.assert_allclose(out[True][0], out[False][0], rtol=2e-5)


@pytest.mark.parametrize('dim', [9, 17, 40, 63])
def test_tc5_narrower_models(dim):
  # 8 < dim < 64: padded factors / assets (zero rows of B, discarded draws)
  from tff_b200 import engine
  dtype = np.float32
  tff, model, (odrift, ovol), x0 = _setup(dim, dtype)
  n, steps = 1500, 7
  kw = dict(num_samples=n, initial_state=x0, num_time_steps=steps)
  payoffs = [engine.european_call(100.0, component=-1), engine.identity(component=dim - 1),
             engine.identity(component=0)]
  with _Tc5(True):
    got = model.price_euler([1.0], payoffs, random_type=tff.math.random.RandomType.SOBOL, **kw)
  with _Tc5(False):
    legacy = model.price_euler([1.0], payoffs, random_type=tff.math.random.RandomType.SOBOL, **kw)
  paths = oeuler.sample(dim, odrift, ovol, [1.0], random_type=odraws.RandomType.SOBOL, dtype=dtype,
                        **kw)[:, 0, :].astype(np.float64)
  want = [np.maximum(paths.mean(axis=1) - 100, 0).mean(), paths[:, dim - 1].mean(), paths[:, 0].mean()]
  np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-4)
  np.testing.assert_allclose(got, legacy, rtol=2e-5, atol=1e-4)


@pytest.mark.parametrize('dim', [17, 64])
def test_tc5_materialised_paths_match_oracle(dim):
  dtype = np.float32
  tff, model, (odrift, ovol), x0 = _setup(dim, dtype)
  kw = dict(num_samples=700, initial_state=x0, skip=300, num_time_steps=10)
  rt = tff.math.random.RandomType.SOBOL
  with _Tc5(True):
    got = model.sample_paths_euler([0.5, 1.0], random_type=rt, **kw).cpu().numpy()
  with _Tc5(False):
    legacy = model.sample_paths_euler([0.5, 1.0], random_type=rt, **kw).cpu().numpy()
  want = oeuler.sample(dim, odrift, ovol, [0.5, 1.0], random_type=odraws.RandomType.SOBOL,
                       dtype=dtype, **kw)
  assert got.shape == want.shape == (700, 2, dim) and got.dtype == dtype
  np.testing.assert_allclose(got, want, rtol=1e-5)
  np.testing.assert_allclose(got, legacy, rtol=1e-5)
