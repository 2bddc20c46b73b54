"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): Sobol points and Philox uint32 streams
bit-exact; normals / path values / prices within 1e-12 relative in float64 and
1e-5 in float32 (a small absolute floor covers values that pass through zero).
"""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import models as omodels
from oracle import philox as ophilox
from oracle import sobol as osobol

pytestmark = pytest.mark.gpu

RTOL = {np.float64: 1e-12, np.float32: 1e-5}
ATOL = {np.float64: 1e-13, np.float32: 2e-6}


def _tff():
  import tff_b200 as tff
  return tff


def _np(t):
  return t.detach().cpu().numpy()


def _close(got, want, dtype, scale=1.0):
  np.testing.assert_allclose(got, want, rtol=RTOL[dtype], atol=ATOL[dtype] * scale)


# ------------------------------------------------------------- Philox ------
@pytest.mark.parametrize('key,ctr,first', [
    ([0, 0], [0, 0, 0, 0], 0),
    ([0xa4093822, 0x299f31d0], [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], 0),
    ([1, 2], [0xfffffff0, 0xffffffff, 0xffffffff, 5], 0),       # carries
    ([7, 9], [0, 0, 3, 4], 2**33 + 5),
])
def test_philox_raw_words_bit_exact(key, ctr, first):
  from tff_b200.math.random import philox
  got = _np(philox.raw_words(key, ctr, first, 4099).view(__import__('torch').int32)).view(np.uint32)
  want = ophilox.raw_words(np.array(key, np.uint32), np.array(ctr, np.uint32), first, 4099)
  np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('shape,seed', [([1001, 7], [4, 2]), ([33], [-5, 123456789012])])
def test_stateless_normal(dtype, shape, seed):
  tff = _tff()
  got = _np(tff.math.random.stateless_normal(shape, seed, dtype))
  want = ophilox.stateless_normal(shape, seed, dtype)
  assert got.dtype == dtype and got.shape == tuple(shape)
  _close(got, want, dtype)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_stateful_normal_first_call(dtype):
  tff = _tff()
  got = _np(tff.math.random.normal([513, 3], dtype=dtype, seed=42))
  want = ophilox.stateful_normal([513, 3], 42, dtype)
  _close(got, want, dtype)


def test_philox_tensorflow_published_vectors_on_the_device():
  # The values TensorFlow prints in its own documentation
  # (tests/test_oracle_kat.py::test_philox_tensorflow_published_*), produced by the
  # CUDA generator through the C ABI.
  tff = _tff()
  from tff_b200.math.random import philox
  got = _np(tff.math.random.stateless_normal([2, 3], [1, 2], np.float32))
  np.testing.assert_allclose(
      got, [[0.5441101, 0.20738031, 0.07356433], [0.04643455, -1.3015898, -0.95385665]],
      rtol=1e-5, atol=1e-7)
  import ctypes as C
  got = _np(philox._fill((C.c_uint32 * 2)(0, 0), (C.c_uint32 * 4)(1, 0, 0, 0),
                         [2, 3], np.float32))
  np.testing.assert_allclose(
      got, [[0.43842277, -0.53439844, -0.07710262], [1.5658046, -0.1012345, -0.2744976]],
      rtol=1e-5, atol=1e-7)


def test_philox_fp64_box_muller_identities_on_the_device():
  # float64: same raw words (bit-exact above), Uint64ToDouble + BoxMullerDouble:
  # z0^2 + z1^2 = -2 ln u1, atan2(z0, z1) = 2 pi u2
  from tff_b200.math.random import philox
  key, ctr = philox.stateless_key_counter([1, 2])
  z = _np(philox._fill(key, ctr, [8192], np.float64)).reshape(-1, 2)
  okey, octr = ophilox.stateless_key_counter([1, 2])
  words = ophilox.raw_words(okey, octr, 0, 4096)
  u1 = np.maximum(ophilox.uint64_to_double(words[:, 0], words[:, 1]), 1e-7)
  u2 = ophilox.uint64_to_double(words[:, 2], words[:, 3])
  np.testing.assert_allclose((z**2).sum(axis=1), -2 * np.log(u1), rtol=1e-12)
  ang = np.arctan2(z[:, 0], z[:, 1]) % (2 * np.pi)
  np.testing.assert_allclose(ang, 2 * np.pi * u2, rtol=0, atol=1e-11)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('first', [1, 2, 3, 1000003])
def test_normal_fill_offsets(dtype, first):
  from tff_b200.math.random import philox
  key, ctr = philox.stateless_key_counter([9, 8])
  got = _np(philox._fill(key, ctr, [777], dtype, first_element=first))
  okey, octr = ophilox.stateless_key_counter([9, 8])
  want = ophilox.normal_fill(okey, octr, 777, dtype, first_element=first)
  _close(got, want, dtype)


# -------------------------------------------------------------- Sobol ------
@pytest.mark.parametrize('dim,n,skip', [(2, 5, 0), (50, 1000, 0), (7, 333, 17),
                                        (1, 3, 2**31 - 5), (600, 64, 123456),
                                        (3, 5000, 2**24 - 100)])
def test_sobol_integer_points_bit_exact(dim, n, skip):
  from tff_b200.math.random import sobol
  got = _np(sobol.sample_integers(dim, n, skip)).astype(np.int64)
  want, _ = osobol.sample_integers(dim, n, skip)
  np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('dim,n,skip', [(2, 5, 0), (40, 2000, 3), (3, 4096, 2**25)])
def test_sobol_uniforms_exact(dtype, dim, n, skip):
  tff = _tff()
  got = _np(tff.math.random.sobol.sample(dim, n, skip=skip, dtype=dtype))
  want = osobol.sample(dim, n, skip=skip, dtype=dtype)
  assert got.dtype == dtype
  np.testing.assert_array_equal(got, want)


def test_sobol_known_values():
  # math/random_ops/sobol/sobol_test.py:28-38 and :93-98 on the device path
  tff = _tff()
  got = _np(tff.math.random.sobol.sample(2, 5, dtype=np.float64))
  np.testing.assert_array_equal(
      got, [[0.5, 0.5], [0.25, 0.75], [0.75, 0.25], [0.125, 0.625], [0.625, 0.125]])
  got = _np(tff.math.random.sobol.sample(1, 3, skip=2**31 - 5))
  np.testing.assert_array_equal(got, [[0.25], [0.75], [0.5]])


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_generate_mc_normal_draws_sobol_kat(dtype):
  # models/utils_test.py:33-55
  tff = _tff()
  got = _np(tff.models.utils.generate_mc_normal_draws(
      num_normal_draws=2, num_time_steps=3, num_sample_paths=4,
      random_type=tff.math.random.RandomType.SOBOL, dtype=dtype, skip=10))
  expected = [[[0.8871465, 0.48877636], [-0.8871465, -0.48877636],
               [0.48877636, 0.8871465], [-0.15731068, 0.15731068]],
              [[0.8871465, -1.5341204], [1.5341204, -0.15731068],
               [-0.15731068, 1.5341204], [-0.8871465, 0.48877636]],
              [[-0.15731068, 1.5341204], [0.15731068, -0.48877636],
               [-1.5341204, 0.8871465], [0.8871465, -1.5341204]]]
  np.testing.assert_allclose(got, expected, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('rt,seed', [('SOBOL', None), ('STATELESS', [1, 2]),
                                     ('STATELESS_ANTITHETIC', [3, 4]),
                                     ('PSEUDO', 11), ('PSEUDO_ANTITHETIC', 12)])
def test_generate_mc_normal_draws_matches_oracle(dtype, rt, seed):
  tff = _tff()
  got = _np(tff.models.utils.generate_mc_normal_draws(
      num_normal_draws=3, num_time_steps=5, num_sample_paths=130,
      random_type=tff.math.random.RandomType[rt], dtype=dtype, seed=seed, skip=7))
  want = odraws.generate_mc_normal_draws(
      3, 5, 130, odraws.RandomType[rt], dtype=dtype, seed=seed, skip=7)
  assert got.shape == want.shape == (5, 130, 3)
  _close(got, want, dtype)


# ------------------------------------------------------------- Euler ------
def _models(dtype):
  tff = _tff()
  from tff_b200.models import closures
  from tff_b200.math import piecewise
  pw = piecewise.PiecewiseConstantFunc([0.3, 0.8], [0.1, 0.2, 0.15], dtype=dtype)
  opw = omodels.PiecewiseConstantFunc([0.3, 0.8], [0.1, 0.2, 0.15], dtype=dtype)
  r, s = 0.03, 0.1
  out = {}
  out['affine_loggbm'] = (
      1, closures.affine_closures(r - s * s / 2, 0.0, s),
      (lambda t, x: (r - s * s / 2) + 0 * x,
       lambda t, x: s * np.ones(x.shape + (1,), dtype=x.dtype)),
      np.array([np.log(700.0)]))
  out['affine_ou'] = (
      1, closures.affine_closures(lambda t: 0.5 * np.sqrt(t), -0.7, lambda t: 0.2 * t + 0.1),
      (lambda t, x: np.asarray(0.5 * np.sqrt(t), x.dtype) - np.asarray(0.7, x.dtype) * x,
       lambda t, x: np.asarray(0.2 * t + 0.1, x.dtype) * np.ones(x.shape + (1,), dtype=x.dtype)),
      np.array([0.1]))
  out['gbm'] = (1, closures.gbm_closures(0.03, pw),
                omodels.gbm_closures(0.03, opw, dtype), np.array([100.0]))
  hm = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=pw,
                              rho=-0.7, dtype=dtype)
  out['heston'] = (2, (hm.drift_fn(), hm.volatility_fn()),
                   omodels.heston_closures(2.0, 0.04, opw, -0.7, dtype),
                   np.array([np.log(100.0), 0.04]))
  return out


GRIDS = [dict(times=[1.0], time_step=0.05),
         dict(times=[0.25, 0.5, 1.0], num_time_steps=12),
         dict(times=[0.0, 0.4, 1.0], time_step=0.1),
         dict(times=[0.3, 0.75], times_grid=[0.0, 0.25, 0.5, 0.75, 1.0, 1.25])]
RNGS = [('SOBOL', None, 0), ('SOBOL', None, 1000), ('STATELESS', [4, 2], 0),
        ('STATELESS_ANTITHETIC', [4, 2], 0), ('PSEUDO', 42, 0),
        ('PSEUDO_ANTITHETIC', 42, 0)]


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('model', ['affine_loggbm', 'affine_ou', 'gbm', 'heston'])
@pytest.mark.parametrize('rng', RNGS, ids=lambda r: f'{r[0]}-{r[2]}')
@pytest.mark.parametrize('gi', range(len(GRIDS)))
def test_euler_paths_match_oracle(dtype, model, rng, gi):
  tff = _tff()
  dim, (drift, vol), (odrift, ovol), x0 = _models(dtype)[model]
  rt, seed, skip = rng
  g = GRIDS[gi]
  n = 1000
  kw = dict(num_samples=n, initial_state=x0.astype(dtype), seed=seed, skip=skip,
            dtype=dtype, **{k: v for k, v in g.items() if k != 'times'})
  got = _np(tff.models.euler_sampling.sample(
      dim, drift, vol, g['times'], random_type=tff.math.random.RandomType[rt], **kw))
  want = oeuler.sample(dim, odrift, ovol, g['times'],
                       random_type=odraws.RandomType[rt], **kw)
  assert got.shape == want.shape == (n, len(g['times']), dim)
  assert got.dtype == dtype
  _close(got, want, dtype, scale=np.abs(want).max())


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_euler_with_supplied_normal_draws(dtype):
  import torch
  tff = _tff()
  dim, (drift, vol), (odrift, ovol), x0 = _models(dtype)['heston']
  rs = np.random.RandomState(0)
  draws = rs.standard_normal((300, 10, 2)).astype(dtype)
  got = _np(tff.models.euler_sampling.sample(
      dim, drift, vol, [0.5, 1.0], time_step=0.1, initial_state=x0.astype(dtype),
      normal_draws=torch.as_tensor(draws).cuda(), dtype=dtype))
  want = oeuler.sample(dim, odrift, ovol, [0.5, 1.0], time_step=0.1,
                       initial_state=x0.astype(dtype), normal_draws=draws, dtype=dtype)
  _close(got, want, dtype, scale=np.abs(want).max())


def test_large_sobol_index_block_alignment():
  # chunks straddle 128-aligned Sobol index blocks with a large skip
  tff = _tff()
  dtype = np.float64
  dim, (drift, vol), (odrift, ovol), x0 = _models(dtype)['heston']
  kw = dict(num_samples=777, initial_state=x0, skip=2**22 + 77, dtype=dtype,
            num_time_steps=6)
  got = _np(tff.models.euler_sampling.sample(
      dim, drift, vol, [1.0], random_type=tff.math.random.RandomType.SOBOL, **kw))
  want = oeuler.sample(dim, odrift, ovol, [1.0],
                       random_type=odraws.RandomType.SOBOL, **kw)
  _close(got, want, dtype, scale=np.abs(want).max())


# ------------------------------------------------------------ pricing ------
@pytest.mark.parametrize('rng', [('SOBOL', None), ('STATELESS', [4, 2]),
                                 ('PSEUDO_ANTITHETIC', 42)], ids=lambda r: r[0])
def test_fused_price_matches_oracle(rng):
  tff = _tff()
  from tff_b200 import engine
  dtype = np.float64
  rt, seed = rng
  dim, (drift, vol), (odrift, ovol), x0 = _models(dtype)['heston']
  n, steps = 4096, 20
  from oracle import grid as ogrid
  all_times, _, _ = ogrid.euler_grid([1.0], dtype=dtype, time_step=0.05)
  assert all_times.shape[0] == steps + 1
  payoffs = [engine.european_call(100.0, log_state=True),
             engine.european_put(95.0, log_state=True, scale=0.97),
             engine.up_and_out_call(100.0, 120.0, log_state=True),
             engine.down_and_out_put(105.0, 85.0, log_state=True),
             engine.identity(component=1)]
  mean, stderr, bad = tff.models.euler_sampling.price(
      dim, drift, vol, [1.0], payoffs, time_step=0.05, num_samples=n,
      initial_state=x0, random_type=tff.math.random.RandomType[rt], seed=seed,
      dtype=dtype, return_stats=True)
  # every grid point of the times=[1.0] run, recorded through times_grid
  paths = oeuler.sample(dim, odrift, ovol, all_times[1:], times_grid=all_times,
                        num_samples=n, initial_state=x0,
                        random_type=odraws.RandomType[rt], seed=seed, dtype=dtype)
  s = np.exp(paths[:, :, 0])
  st = s[:, -1]
  smax = np.maximum(s.max(axis=1), 100.0)
  smin = np.minimum(s.min(axis=1), 100.0)
  want = [np.maximum(st - 100, 0),
          0.97 * np.maximum(95 - st, 0),
          np.where(smax > 120.0, 0.0, np.maximum(st - 100, 0)),
          np.where(smin < 85.0, 0.0, np.maximum(105 - st, 0)),
          paths[:, -1, 1]]
  np.testing.assert_allclose(mean, [w.mean() for w in want], rtol=1e-12)
  np.testing.assert_allclose(
      stderr, [np.sqrt(max((w**2).mean() - w.mean()**2, 0) / n) for w in want],
      rtol=1e-9)
  assert np.all(bad == 0)


def test_sharded_price_adds_up():
  tff = _tff()
  from tff_b200 import engine
  from tff_b200.models import closures
  dtype = np.float64
  dim, (drift, vol), _, x0 = _models(dtype)['heston']
  spec = closures.resolve_spec(drift, vol)
  times = np.linspace(0, 1, 11)
  pay = [engine.european_call(100.0, log_state=True)]
  for rt, seed in (('SOBOL', None), ('STATELESS', [1, 2]), ('STATELESS_ANTITHETIC', [1, 2])):
    rng = engine.RngSpec(tff.math.random.RandomType[rt], seed, 5)
    plan = engine.Plan(spec, times, 10, x0, rng, 3000, dtype)
    full = _np(plan.price_sums(pay))
    units = plan.units
    parts = [_np(plan.price_sums(pay, a, b - a))
             for a, b in ((0, 1001), (1001, 1002), (1002, units))]
    np.testing.assert_allclose(sum(parts)[:, :2], full[:, :2], rtol=1e-13)
    plan.close()


def test_c1_notebook_configuration():
  # examples/jupyter_notebooks/Monte_Carlo_Euler_Scheme.ipynb:223-275
  tff = _tff()
  from tff_b200 import engine
  from tff_b200.models import closures
  r, sigma, spot = 0.03, 0.1, 700.0
  strikes = [600.0, 650.0, 680.0]
  drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
  process = tff.models.GenericItoProcess(1, drift, vol, dtype=np.float64)
  n = 100000
  payoffs = [engine.european_call(k, log_state=True, scale=np.exp(-r)) for k in strikes]
  got = process.price([1.0], payoffs, num_samples=n,
                      initial_state=np.array([np.log(spot)]),
                      random_type=tff.math.random.RandomType.PSEUDO_ANTITHETIC,
                      seed=42, time_step=0.01)
  ref = oeuler.sample(
      1, lambda t, x: (r - sigma**2 / 2) + 0 * x,
      lambda t, x: sigma * np.ones(x.shape + (1,)), [1.0], time_step=0.01,
      num_samples=n, initial_state=np.array([np.log(spot)]),
      random_type=odraws.RandomType.PSEUDO_ANTITHETIC, seed=42, dtype=np.float64)
  want = [np.exp(-r) * np.maximum(np.exp(ref[:, 0, 0]) - k, 0).mean() for k in strikes]
  np.testing.assert_allclose(got, want, rtol=1e-12)
  # Black-Scholes sanity (statistical, 3 standard errors)
  from scipy.stats import norm
  for k, g in zip(strikes, got):
    d1 = (np.log(spot / k) + (r + sigma**2 / 2)) / sigma
    bs = spot * norm.cdf(d1) - k * np.exp(-r) * norm.cdf(d1 - sigma)
    assert abs(g - bs) < 0.5


# ---------------------------------------------------------- Heston QE ------
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('rng', [('SOBOL', None, 0), ('STATELESS', [4, 2], 0),
                                 ('PSEUDO_ANTITHETIC', 3, 0)], ids=lambda r: r[0])
@pytest.mark.parametrize('grid', [dict(time_step=0.05), dict(num_time_steps=7),
                                  dict(times_grid=[0.0, 0.1, 0.3, 0.5, 0.8, 1.0, 1.3])],
                         ids=['dt', 'nsteps', 'grid'])
def test_heston_qe_matches_oracle(dtype, rng, grid):
  # HestonModel.sample_paths = Andersen QE (heston_model.py:177-460)
  from oracle import heston_qe as oqe
  tff = _tff()
  from tff_b200.math import piecewise
  rt, seed, skip = rng
  volvol = piecewise.PiecewiseConstantFunc([0.5], [0.5, 0.9], dtype=dtype)
  ovolvol = omodels.PiecewiseConstantFunc([0.5], [0.5, 0.9], dtype=dtype)
  model = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=volvol,
                                 rho=-0.7, dtype=dtype)
  x0 = np.array([np.log(100.0), 0.02], dtype=dtype)
  times = [0.5, 1.0]
  n = 2000
  got = _np(model.sample_paths(times, x0, num_samples=n,
                               random_type=tff.math.random.RandomType[rt], seed=seed,
                               skip=skip, **grid))
  want = oqe.sample_paths(2.0, 0.04, ovolvol, -0.7, times, x0, num_samples=n,
                          random_type=odraws.RandomType[rt], seed=seed, skip=skip,
                          dtype=dtype, **grid)
  assert got.shape == want.shape == (n, 2, 2) and got.dtype == dtype
  if dtype == np.float64:
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
  else:
    # float32: the psi < 1.5 switch can flip for isolated paths; compare robustly
    close = np.isclose(got, want, rtol=2e-4, atol=2e-5)
    assert close.mean() > 0.999


# ------------------------------------------------ exact GBM samplers ------
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('rng', [('SOBOL', None, 5), ('STATELESS', [4, 2], 0),
                                 ('STATELESS_ANTITHETIC', [1, 2], 0)], ids=lambda r: r[0])
def test_exact_gbm_sample_paths(dtype, rng):
  # GeometricBrownianMotion.sample_paths (univariate_...py:155-317)
  tff = _tff()
  from tff_b200.math import piecewise
  rt, seed, skip = rng
  vol = piecewise.PiecewiseConstantFunc([0.3, 0.8], [0.1, 0.2, 0.15], dtype=dtype)
  ovol = omodels.PiecewiseConstantFunc([0.3, 0.8], [0.1, 0.2, 0.15], dtype=dtype)
  model = tff.models.GeometricBrownianMotion(0.03, vol, dtype=dtype)
  times = [0.1, 0.5, 1.0, 2.0]
  n = 3000
  got = _np(model.sample_paths(times, initial_state=np.array([1.5], dtype), num_samples=n,
                               random_type=tff.math.random.RandomType[rt], seed=seed, skip=skip))
  want = omodels.gbm_exact_sample_paths(0.03, ovol, times, np.array([1.5], dtype), n,
                                        odraws.RandomType[rt], seed, skip, dtype)
  assert got.shape == want.shape == (n, 4, 1) and got.dtype == dtype
  np.testing.assert_allclose(got, want, rtol=1e-12 if dtype == np.float64 else 2e-5)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_exact_multivariate_gbm_sample_paths(dtype):
  # MultivariateGeometricBrownianMotion.sample_paths (multivariate_...py:153-282)
  tff = _tff()
  dim = 5
  means = np.linspace(0.01, 0.05, dim).astype(dtype)
  vols = np.linspace(0.1, 0.3, dim).astype(dtype)
  corr = (0.2 + 0.8 * np.eye(dim)).astype(dtype)
  model = tff.models.MultivariateGeometricBrownianMotion(dim, means, vols, corr, dtype=dtype)
  times = [0.25, 1.0, 1.5]
  x0 = np.linspace(1.0, 2.0, dim).astype(dtype)
  n = 2500
  got = _np(model.sample_paths(times, initial_state=x0, num_samples=n,
                               random_type=tff.math.random.RandomType.SOBOL, skip=11))
  want = omodels.mvgbm_exact_sample_paths(means, vols, corr, times, x0, n,
                                          odraws.RandomType.SOBOL, None, 11, dtype)
  assert got.shape == want.shape == (n, 3, dim)
  np.testing.assert_allclose(got, want, rtol=1e-12 if dtype == np.float64 else 2e-5)


# ------------------------------------------------------------ edge cases ----
def test_many_steps_tables_in_global_memory():
  # 15 000 steps: the coefficient table (720 KB) no longer fits shared memory
  tff = _tff()
  dtype = np.float64
  dim, (drift, vol), (odrift, ovol), x0 = _models(dtype)['affine_ou']
  kw = dict(num_samples=256, initial_state=x0, seed=[4, 2], time_step=1.0 / 15000, dtype=dtype)
  got = _np(tff.models.euler_sampling.sample(
      dim, drift, vol, [0.5, 1.0], random_type=tff.math.random.RandomType.STATELESS, **kw))
  want = oeuler.sample(dim, odrift, ovol, [0.5, 1.0],
                       random_type=odraws.RandomType.STATELESS, **kw)
  np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('n', [1, 2, 127, 128, 129, 1025])
def test_ragged_path_counts(n):
  tff = _tff()
  dtype = np.float64
  dim, (drift, vol), (odrift, ovol), x0 = _models(dtype)['heston']
  for rt, seed, skip in (('SOBOL', None, 126), ('STATELESS', [1, 1], 0)):
    kw = dict(num_samples=n, initial_state=x0, seed=seed, skip=skip, num_time_steps=5, dtype=dtype)
    got = _np(tff.models.euler_sampling.sample(
        dim, drift, vol, [1.0], random_type=tff.math.random.RandomType[rt], **kw))
    want = oeuler.sample(dim, odrift, ovol, [1.0], random_type=odraws.RandomType[rt], **kw)
    assert got.shape == (n, 1, 2)
    _close(got, want, dtype, scale=np.abs(want).max())


def test_only_initial_time_requested():
  tff = _tff()
  dtype = np.float64
  dim, (drift, vol), _, x0 = _models(dtype)['heston']
  got = _np(tff.models.euler_sampling.sample(
      dim, drift, vol, [0.0], num_samples=10, initial_state=x0, time_step=0.1, seed=[1, 2],
      random_type=tff.math.random.RandomType.STATELESS, dtype=dtype))
  np.testing.assert_array_equal(got, np.broadcast_to(x0, (10, 1, 2)))


def test_limits_are_reported():
  tff = _tff()
  dtype = np.float64
  dim, (drift, vol), _, x0 = _models(dtype)['heston']
  sample = tff.models.euler_sampling.sample
  rt = tff.math.random.RandomType
  with pytest.raises(ValueError):          # Sobol dimension 2 * 11000 > 21201
    sample(dim, drift, vol, [1.0], num_samples=8, initial_state=x0, num_time_steps=11000,
           random_type=rt.SOBOL, dtype=dtype)
  with pytest.raises(ValueError):          # skip + num_samples >= 2^31 - 1
    sample(dim, drift, vol, [1.0], num_samples=8, initial_state=x0, num_time_steps=4,
           random_type=rt.SOBOL, skip=2**31 - 5, dtype=dtype)
  with pytest.raises(ValueError):          # STATELESS needs a seed
    sample(dim, drift, vol, [1.0], num_samples=8, initial_state=x0, num_time_steps=4,
           random_type=rt.STATELESS, dtype=dtype)
  with pytest.raises(ValueError):          # per-path initial states: one row per sample
    sample(dim, drift, vol, [1.0], num_samples=6,
           initial_state=np.arange(16.0).reshape(8, 2), num_time_steps=4, seed=1, dtype=dtype)


def test_pseudo_without_seed_is_random_but_valid():
  tff = _tff()
  dtype = np.float64
  dim, (drift, vol), _, x0 = _models(dtype)['gbm']
  a = _np(tff.models.euler_sampling.sample(dim, drift, vol, [1.0], num_samples=4096,
                                           initial_state=x0, time_step=0.1, dtype=dtype))
  b = _np(tff.models.euler_sampling.sample(dim, drift, vol, [1.0], num_samples=4096,
                                           initial_state=x0, time_step=0.1, dtype=dtype))
  assert np.all(np.isfinite(a)) and not np.array_equal(a, b)
  assert abs(np.log(a / 100.0).mean() - (0.03 - 0.5 * 0.0225)) < 0.02


# ------------------------------------------------------ batched processes ----
@pytest.mark.parametrize('rng', [('SOBOL', None, 3), ('STATELESS', [4, 2], 0),
                                 ('STATELESS_ANTITHETIC', [4, 2], 0), ('PSEUDO_ANTITHETIC', 9, 0)],
                         ids=lambda r: r[0])
def test_batch_of_initial_states(rng):
  # batch_shape = initial_state.shape[:-2] (euler_sampling.py:251); draws laid
  # out [steps] + batch + [N, dim] (models/utils.py:98-128)
  tff = _tff()
  dtype = np.float64
  rt, seed, skip = rng
  dim, (drift, vol), (odrift, ovol), _ = _models(dtype)['heston']
  x0 = np.array([[[np.log(100.0), 0.04]], [[np.log(90.0), 0.09]], [[np.log(120.0), 0.01]]])
  n = 512
  kw = dict(num_samples=n, initial_state=x0, seed=seed, skip=skip, num_time_steps=6, dtype=dtype)
  got = _np(tff.models.euler_sampling.sample(
      dim, drift, vol, [0.5, 1.0], random_type=tff.math.random.RandomType[rt], **kw))
  want = oeuler.sample(dim, odrift, ovol, [0.5, 1.0], random_type=odraws.RandomType[rt], **kw)
  assert got.shape == want.shape == (3, n, 2, 2)
  _close(got, want, dtype, scale=np.abs(want).max())


def test_batch_of_gbm_parameters():
  # batched GBM parameters of shape batch_shape + [1] (univariate_...py:66-80)
  tff = _tff()
  from tff_b200.models import closures
  dtype = np.float64
  mean = np.array([[0.01], [0.05]])
  vol = np.array([[0.1], [0.3]])
  drift, volf = closures.gbm_closures(mean, vol)
  x0 = np.array([[[1.0]], [[2.0]]])
  n = 700
  kw = dict(num_samples=n, initial_state=x0, seed=[1, 5], time_step=0.1, dtype=dtype)
  got = _np(tff.models.euler_sampling.sample(
      1, drift, volf, [1.0], random_type=tff.math.random.RandomType.STATELESS, **kw))
  want = oeuler.sample(1, lambda t, x: mean[:, None, :] * x,
                       lambda t, x: (vol[:, None, :] * x)[..., None], [1.0],
                       random_type=odraws.RandomType.STATELESS, **kw)
  assert got.shape == want.shape == (2, n, 1, 1)
  _close(got, want, dtype, scale=np.abs(want).max())


# ----- tff.math.random.uniform (math/random_ops/uniform.py; uniform_test.py:31-73)
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_uniform_stateless_and_pseudo_bit_exact(dtype):
  tff = _tff()
  rt = tff.math.random.RandomType
  got = _np(tff.math.random.uniform(10, [2, 3, 5000], random_type=rt.STATELESS, seed=[2, 2], dtype=dtype))
  want = ophilox.stateless_uniform([2, 3, 5000, 10], [2, 2], dtype)
  assert got.dtype == dtype and got.shape == (2, 3, 5000, 10)
  np.testing.assert_array_equal(got, want)          # pure bit manipulation + one exact subtraction
  assert got.min() >= 0.0 and got.max() < 1.0
  np.testing.assert_array_almost_equal(got.mean(axis=2), 0.5 * np.ones((2, 3, 10)), decimal=2)
  got = _np(tff.math.random.uniform(10, [2, 3, 5000], seed=101, dtype=dtype))
  np.testing.assert_array_equal(got, ophilox.stateful_uniform([2, 3, 5000, 10], 101, dtype))
  np.testing.assert_array_almost_equal(got.mean(axis=2), 0.5 * np.ones((2, 3, 10)), decimal=2)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_uniform_sobol_equals_sobol_sample(dtype):
  # uniform_test.py:58-73
  tff = _tff()
  got = _np(tff.math.random.uniform(5, [100], random_type=tff.math.random.RandomType.SOBOL,
                                    skip=1000, dtype=dtype))
  want = _np(tff.math.random.sobol.sample(dim=5, num_results=100, skip=1000, dtype=dtype))
  assert got.dtype == dtype
  np.testing.assert_array_equal(got, want)


def test_uniform_argument_errors():
  tff = _tff()
  rt = tff.math.random.RandomType
  with pytest.raises(ValueError):
    tff.math.random.uniform(2, [4], random_type=rt.STATELESS)
  with pytest.raises(NotImplementedError):
    tff.math.random.uniform(2, [4], random_type=rt.PSEUDO_ANTITHETIC, seed=1)


# ----- stateless_random_shuffle (math/random_ops/stateless.py:24-52; stateless_test.py:29-125)
def test_stateless_random_shuffle_matches_reference_construction():
  import torch
  tff = _tff()
  shuffles = {}
  for dtype in (np.int32, np.int64, np.float32, np.float64):
    ident = np.arange(10, dtype=dtype)
    s1 = _np(tff.math.random.stateless_random_shuffle(ident, seed=(1, 42)))
    s2 = _np(tff.math.random.stateless_random_shuffle(ident, seed=(2, 42)))
    assert s1.dtype == dtype and s2.dtype == dtype
    assert np.abs(s1 - s2).max() > 0                               # different seeds differ
    assert set(s1.tolist()) == set(ident.tolist()) == set(s2.tolist())   # permutations
    # the permutation is argsort(stable) of the float64 stateless uniforms
    want = ident[np.argsort(ophilox.stateless_uniform([10], [1, 42], np.float64), kind='stable')]
    np.testing.assert_array_equal(s1, want)
    shuffles[dtype] = _np(tff.math.random.stateless_random_shuffle(ident, seed=(100, 42)))
    np.testing.assert_array_equal(                                  # stateless
        shuffles[dtype], _np(tff.math.random.stateless_random_shuffle(ident, seed=(100, 42))))
  for dtype in shuffles:                                            # same across dtypes
    np.testing.assert_array_equal(shuffles[dtype], shuffles[np.int32].astype(dtype))
  # independent of the input values
  rs = np.random.RandomState(25)
  random_input = np.sort(rs.normal(size=[10]))
  control = _np(tff.math.random.stateless_random_shuffle(random_input, seed=(100, 42)))
  np.testing.assert_array_equal(np.argsort(shuffles[np.int64]), np.argsort(control))
  # multi-dimensional input: rows are shuffled
  x = np.array([[[1], [2], [3]], [[4], [5], [6]]], dtype=np.float32)
  got = _np(tff.math.random.stateless_random_shuffle(torch.as_tensor(x), seed=(1, 42)))
  assert got.shape == x.shape and got.dtype == x.dtype
  assert sorted(got.reshape(2, 3).tolist()) == sorted(x.reshape(2, 3).tolist())


# ----- one initial state per path (`initial_state` of shape [num_samples, dim],
# euler_sampling.py:357: `initial_state + zeros([num_samples, dim])`)
@pytest.mark.parametrize('rt', ['STATELESS_ANTITHETIC', 'SOBOL', 'STATELESS'])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_per_path_initial_states(rt, dtype):
  tff = _tff()
  from tff_b200.models import closures
  from oracle import euler as oeuler
  n = 3002
  rs = np.random.RandomState(3)
  x0 = rs.uniform(0.5, 2.0, size=(n, 1)).astype(dtype)
  mu, sigma = 0.05, 0.3
  drift, vol = closures.gbm_closures(mu, sigma)
  kw = dict(num_samples=n, initial_state=x0, seed=[4, 2], time_step=0.05, dtype=dtype)
  got = _np(tff.models.euler_sampling.sample(1, drift, vol, [0.0, 0.5, 1.0],
                                             random_type=tff.math.random.RandomType[rt], **kw))
  want = oeuler.sample(1, lambda t, x: dtype(mu) * x, lambda t, x: (dtype(sigma) * x)[..., None],
                       [0.0, 0.5, 1.0], random_type=odraws.RandomType[rt], **kw)
  assert got.shape == want.shape == (n, 3, 1)
  np.testing.assert_array_equal(got[:, 0, :], x0)
  _close(got, want, dtype)
  # Heston (dim 2), per-path spot and variance
  x02 = np.stack([rs.uniform(4.0, 5.0, n), rs.uniform(0.02, 0.08, n)], -1).astype(dtype)
  heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=dtype)
  got = _np(heston.sample_paths_euler([0.5, 1.0], x02, num_samples=n, num_time_steps=20,
                                      random_type=tff.math.random.RandomType[rt], seed=[4, 2]))
  from oracle import models as omodels
  od, ov = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, dtype)
  want = oeuler.sample(2, od, ov, [0.5, 1.0], num_samples=n, initial_state=x02, num_time_steps=20,
                       random_type=odraws.RandomType[rt], seed=[4, 2], dtype=dtype)
  _close(got, want, dtype)
  with pytest.raises(ValueError):
    tff.models.euler_sampling.sample(1, drift, vol, [1.0], num_samples=n + 2, initial_state=x0,
                                     seed=[4, 2], time_step=0.05, dtype=dtype,
                                     random_type=tff.math.random.RandomType[rt])
