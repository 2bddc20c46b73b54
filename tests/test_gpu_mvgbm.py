"""GPU parity of the correlated multi-asset GBM Euler kernel (config C4 shape)
against the oracle; float32 tolerance 1e-5, float64 1e-12."""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import models as omodels

pytestmark = pytest.mark.gpu


def _setup(dim, dtype):
  import tff_b200 as tff
  means = np.full(dim, 0.03, dtype=dtype)
  vols = np.linspace(0.1, 0.4, dim).astype(dtype)
  corr = (0.3 + 0.7 * np.eye(dim)).astype(dtype)
  model = tff.models.MultivariateGeometricBrownianMotion(
      dim, means=means, volatilities=vols, corr_matrix=corr, dtype=dtype)
  oclos = omodels.mvgbm_closures(means, vols, corr, dtype)
  x0 = (100.0 * np.ones(dim)).astype(dtype)
  return tff, model, oclos, x0


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('dim', [3, 8, 17, 64])
@pytest.mark.parametrize('rng', [('SOBOL', None, 0), ('SOBOL', None, 300),
                                 ('STATELESS', [4, 2], 0)], ids=lambda r: f'{r[0]}{r[2]}')
def test_paths_match_oracle(dtype, dim, rng):
  tff, model, (odrift, ovol), x0 = _setup(dim, dtype)
  rt, seed, skip = rng
  n = 600 if dim == 64 else 1500
  kw = dict(num_samples=n, initial_state=x0, seed=seed, skip=skip, num_time_steps=10)
  got = model.sample_paths_euler([0.5, 1.0], random_type=tff.math.random.RandomType[rt],
                                 **kw).cpu().numpy()
  want = oeuler.sample(dim, odrift, ovol, [0.5, 1.0], random_type=odraws.RandomType[rt],
                       dtype=dtype, **kw)
  assert got.shape == want.shape == (n, 2, dim) and got.dtype == dtype
  if dtype == np.float32:
    np.testing.assert_allclose(got, want, rtol=1e-5)
  else:
    np.testing.assert_allclose(got, want, rtol=1e-12)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_basket_price_matches_oracle(dtype):
  from tff_b200 import engine
  dim = 64
  tff, model, (odrift, ovol), x0 = _setup(dim, dtype)
  n = 4096
  kw = dict(num_samples=n, initial_state=x0, num_time_steps=12)
  got = model.price_euler([1.0], [engine.european_call(100.0, component=-1),
                                  engine.european_put(100.0, component=-1),
                                  engine.identity(component=5)],
                          random_type=tff.math.random.RandomType.SOBOL, **kw)
  paths = oeuler.sample(dim, odrift, ovol, [1.0], random_type=odraws.RandomType.SOBOL,
                        dtype=dtype, **kw)[:, 0, :].astype(np.float64)
  basket = paths.mean(axis=1)
  want = [np.maximum(basket - 100, 0).mean(), np.maximum(100 - basket, 0).mean(),
          paths[:, 5].mean()]
  np.testing.assert_allclose(got, want, rtol=2e-5 if dtype == np.float32 else 1e-12)


def test_c4_shape_252_steps_has_no_systematic_error():
  # the C4 grid (252 steps, 64 assets, float32, Sobol): elementwise within the
  # float32 tolerance AND the per-asset means within 2e-6 -- a rounding error
  # common to all paths and steps (e.g. forming 1 + mu dt in float32) passes a
  # 10-step test and shows up here as a 1e-4 shift of the means
  dtype, dim, n = np.float32, 64, 512
  tff, model, (odrift, ovol), x0 = _setup(dim, dtype)
  kw = dict(num_samples=n, initial_state=x0, num_time_steps=252)
  got = model.sample_paths_euler([1.0], random_type=tff.math.random.RandomType.SOBOL,
                                 **kw).cpu().numpy()[:, 0, :].astype(np.float64)
  want = oeuler.sample(dim, odrift, ovol, [1.0], random_type=odraws.RandomType.SOBOL,
                       dtype=dtype, **kw)[:, 0, :].astype(np.float64)
  np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6 * 100)
  np.testing.assert_allclose(got.mean(axis=0), want.mean(axis=0), rtol=2e-6)


def test_fp32_sobol_uniform_equal_to_one_gives_inf():
  # beyond 2^24 points a float32 Sobol uniform can round to 1.0 and
  # the reference's erfinv returns +inf; the engine reproduces and counts it.
  import tff_b200 as tff
  from tff_b200 import engine
  tff_, model, _, x0 = _setup(16, np.float32)
  # dimension 12 hits u == 1.0 at point 18 684 944 (index = skip + 1 + p)
  mean, stderr, bad = model.price_euler(
      [1.0], [engine.identity(component=12)], initial_state=x0, num_samples=256,
      random_type=tff.math.random.RandomType.SOBOL, skip=18684944 - 100,
      num_time_steps=1, return_stats=True)
  assert bad[0] >= 1


def test_fp32_sobol_clamped_mode_keeps_the_hazard_paths_finite():
  # tqf_plan_set_sobol_clamp: u == 1.0 -> largest float32 below one (z = ndtri(1 - 2^-24)
  # = 5.42); every other draw, hence every other path, is unchanged.
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  from tff_b200.models import utils
  tff_, model, _, x0 = _setup(16, np.float32)
  spec = closures.resolve_spec(model.drift_fn(), model.volatility_fn())
  all_times, mask, _ = utils.prepare_grid(times=np.array([1.0], np.float32),
                                          time_step=np.float32(1.0), num_time_steps=1,
                                          dtype=np.float32)
  steps, _ = engine.record_plan(mask, 1)
  rng = engine.RngSpec(tff.math.random.RandomType.SOBOL, None, 18684944 - 100)
  plan = engine.Plan(spec, all_times, steps, x0, rng, 256, np.float32)
  try:
    pay = [engine.identity(component=12), engine.european_call(100.0, component=-1)]
    strict = plan.price_sums(pay).cpu().numpy()
    plan.set_sobol_clamp(True)
    clamped = plan.price_sums(pay).cpu().numpy()
    plan.set_sobol_clamp(False)
    again = plan.price_sums(pay).cpu().numpy()
  finally:
    plan.close()
  assert strict[0, 2] >= 1 and np.all(clamped[:, 2] == 0)
  np.testing.assert_array_equal(strict, again)
  # the hazard path now contributes a finite value: asset 12 moved by +5.42 sigma sqrt(dt)
  extra = clamped[0, 0] - strict[0, 0]
  assert 100.0 < extra / strict[0, 2] < 2000.0
