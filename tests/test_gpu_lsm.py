"""GPU parity of the Longstaff-Schwartz passes: the reference's own KATs
(models/longstaff_schwartz/lsm_test.py) and the oracle on simulated paths."""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import lsm as olsm

pytestmark = pytest.mark.gpu

_SAMPLES = np.expand_dims([[1.0, 1.09, 1.08, 1.34], [1.0, 1.16, 1.26, 1.54],
                           [1.0, 1.22, 1.07, 1.03], [1.0, 0.93, 0.97, 0.92],
                           [1.0, 1.11, 1.56, 1.52], [1.0, 0.76, 0.77, 0.90],
                           [1.0, 0.92, 0.84, 1.01], [1.0, 0.88, 1.22, 1.34]], -1)
_DF = np.exp(-np.cumsum([0.06, 0.06, 0.06]))


def _lsm():
  import tff_b200 as tff
  return tff.models.longstaff_schwartz


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_reference_kats(dtype):
  # lsm_test.py:60-127, 169-219
  lsm = _lsm()
  basis = lsm.make_polynomial_basis(2)
  put = lsm.make_basket_put_payoff([1.1], dtype=dtype)
  tol = dict(rtol=1e-4, atol=1e-4)
  got = lsm.least_square_mc(_SAMPLES, [3], put, basis, discount_factors=[_DF[-1]], dtype=dtype)
  assert got.dtype == dtype and got.shape == (1,)
  np.testing.assert_allclose(got, [0.0564], **tol)
  np.testing.assert_allclose(
      lsm.least_square_mc(_SAMPLES, [1, 2, 3], put, basis, discount_factors=_DF, dtype=dtype),
      [0.1144], **tol)
  np.testing.assert_allclose(
      lsm.least_square_mc(_SAMPLES, [1, 2, 3], put, basis, discount_factors=_DF,
                          num_calibration_samples=4, dtype=dtype), [0.174226], **tol)
  put2 = lsm.make_basket_put_payoff([1.1, 1.2], dtype=dtype)
  df2 = np.exp(-np.cumsum([[0.06] * 3, [0.05] * 3], -1))[None]
  np.testing.assert_allclose(
      lsm.least_square_mc(_SAMPLES, [1, 2, 3], put2, basis, discount_factors=df2, dtype=dtype),
      [0.1144, 0.199], **tol)
  batch = np.stack([_SAMPLES, _SAMPLES + 0.1], 0)
  np.testing.assert_allclose(
      lsm.least_square_mc(batch, [1, 2, 3], put2, basis, discount_factors=df2, dtype=dtype),
      [0.1144, 0.1157], **tol)


def test_basket_degree_10_generic_path():
  # lsm_test.py:129-157: K = 121 basis functions on 8 samples (rank deficient)
  lsm = _lsm()
  basis = lsm.make_polynomial_basis(10)
  put = lsm.make_basket_put_payoff([1.1, 1.2, 1.3], dtype=np.float64)
  s2 = np.concatenate([_SAMPLES, _SAMPLES], -1)
  a = lsm.least_square_mc(s2, [1, 2, 3], put, basis, discount_factors=_DF, dtype=np.float64)
  b = lsm.least_square_mc(_SAMPLES, [1, 2, 3], put, basis, discount_factors=_DF, dtype=np.float64)
  assert a.shape == (3,)
  np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)
  ob = olsm.least_square_mc(_SAMPLES, [1, 2, 3], olsm.make_basket_put_payoff([1.1, 1.2, 1.3]),
                            olsm.make_polynomial_basis(10), _DF, dtype=np.float64)
  np.testing.assert_allclose(b, ob, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('degree,calib', [(3, None), (2, 20000), (5, None)])
def test_american_put_on_engine_paths_matches_oracle(degree, calib):
  # the docstring example of lsm.py:145-183 on the engine's own paths
  import torch
  import tff_b200 as tff
  from tff_b200.models import closures
  lsm = _lsm()
  r, sigma, n = 0.1, 1.0, 1 << 16
  times = np.linspace(0.0, 1.0, 13)
  drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
  kw = dict(num_samples=n, initial_state=np.array([0.0]), seed=[4, 2], time_step=0.05,
            dtype=np.float64)
  log_paths = tff.models.euler_sampling.sample(
      1, drift, vol, times, random_type=tff.math.random.RandomType.STATELESS_ANTITHETIC, **kw)
  paths = torch.exp(log_paths)                       # keeps the time-major strides
  df = np.exp(-r * times)
  got = lsm.least_square_mc(paths, np.arange(13), lsm.make_basket_put_payoff([1.1], dtype=np.float64),
                            lsm.make_polynomial_basis(degree), discount_factors=df,
                            num_calibration_samples=calib, dtype=np.float64)
  opaths = np.exp(oeuler.sample(
      1, lambda t, x: (r - sigma**2 / 2) + 0 * x, lambda t, x: sigma * np.ones(x.shape + (1,)),
      times, random_type=odraws.RandomType.STATELESS_ANTITHETIC, **kw))
  np.testing.assert_allclose(paths.cpu().numpy(), opaths, rtol=1e-12)
  want = olsm.least_square_mc(opaths, np.arange(13), olsm.make_basket_put_payoff([1.1]),
                              olsm.make_polynomial_basis(degree), df,
                              num_calibration_samples=calib, dtype=np.float64)
  # the regression is solved from differently ordered sums: exercise decisions
  # of paths within ~1e-9 of the boundary may flip -> 1e-9 relative on the price
  np.testing.assert_allclose(got, want, rtol=1e-9)


def test_two_dimensional_basket_matches_oracle():
  lsm = _lsm()
  rs = np.random.RandomState(5)
  n, T = 5000, 6
  steps = rs.standard_normal((n, T, 2)) * 0.1
  paths = np.exp(np.cumsum(steps, axis=1))
  df = np.exp(-0.05 * np.arange(1, T + 1))
  got = lsm.least_square_mc(paths, np.arange(T), lsm.make_basket_put_payoff([1.0, 1.1], dtype=np.float64),
                            lsm.make_polynomial_basis(2), discount_factors=df, dtype=np.float64)
  want = olsm.least_square_mc(paths, np.arange(T), olsm.make_basket_put_payoff([1.0, 1.1]),
                              olsm.make_polynomial_basis(2), df, dtype=np.float64)
  np.testing.assert_allclose(got, want, rtol=1e-9)


def test_lsm_tabulated_payoff_and_per_path_discounting():
  # the shape of the Bermudan swaption problem (hull_white/swaption.py:608-724):
  # exercise values precomputed per (date, path, payoff), one discount curve per
  # path, quadratic basis on a 1-d state
  import torch
  import tff_b200 as tff
  from oracle import lsm as olsm
  rs = np.random.RandomState(11)
  n, t, b = 20000, 6, 3
  state = np.cumsum(0.01 * rs.standard_normal((n, t, 1)), axis=1) + 0.02       # short rate
  df = np.exp(-np.cumsum(np.abs(state[:, :, 0]) * 0.5, axis=1))[:, None, :]     # [N, 1, T]
  values = np.maximum(rs.standard_normal((t, n, b)) * 0.01 +
                      (state[:, :, 0].T[..., None] - 0.02) * np.array([1.0, -1.0, 0.5]), 0.0)
  values[2, :, 1] = 0.0                       # a date on which payoff 1 cannot be exercised
  lsm = tff.models.longstaff_schwartz
  got = lsm.least_square_mc(
      torch.as_tensor(state).cuda(), np.arange(t), lsm.make_tabulated_payoff(torch.as_tensor(values).cuda()),
      lsm.make_polynomial_basis(2), discount_factors=torch.as_tensor(df).cuda(), dtype=np.float64)
  want = olsm.least_square_mc(
      state, np.arange(t), lambda x, ti: values[ti], olsm.make_polynomial_basis(2),
      discount_factors=df, dtype=np.float64)
  assert got.shape == (b,)
  np.testing.assert_allclose(got, want, rtol=1e-10)


def test_column_sums_from_the_path_kernel():
  # engine.Plan.paths(column_sums=True): the sums the LSM basis means need,
  # accumulated by the kernel that stores the paths (lsm.py:110-111)
  import torch
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures, utils
  lsm = _lsm()
  r, sigma, n = 0.1, 1.0, 50_002
  times = np.linspace(0.0, 1.0, 13)
  drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
  spec = closures.resolve_spec(drift, vol)
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.05), dtype=np.float64)
  steps, record_slot = engine.record_plan(mask, 13)
  for rtype in (tff.math.random.RandomType.STATELESS_ANTITHETIC, tff.math.random.RandomType.SOBOL):
    rng = engine.RngSpec(rtype, [4, 2], 0)
    plan = engine.Plan(spec, all_times, steps, np.array([0.0]), rng, n, np.float64)
    try:
      ref = plan.paths(record_slot, 13, exp_transform=True)
      paths, sums = plan.paths(record_slot, 13, exp_transform=True, column_sums=True)
      assert torch.equal(ref, paths)
      np.testing.assert_allclose(sums.cpu().numpy(), paths.sum(dim=0).cpu().numpy(), rtol=1e-13)
      # a shard of the units
      p2, s2 = plan.paths(record_slot, 13, 1000, 3000, exp_transform=True, column_sums=True)
      np.testing.assert_allclose(s2.cpu().numpy(), p2.sum(dim=0).cpu().numpy(), rtol=1e-13)
      df = np.exp(-r * times)
      put, basis = lsm.make_basket_put_payoff([1.1], dtype=np.float64), lsm.make_polynomial_basis(3)
      a = lsm.least_square_mc(paths, np.arange(13), put, basis, discount_factors=df, dtype=np.float64)
      b = lsm.least_square_mc(paths, np.arange(13), put, basis, discount_factors=df, dtype=np.float64,
                              column_sums=sums)
      np.testing.assert_allclose(a, b, rtol=1e-12)
    finally:
      plan.close()


@pytest.mark.parametrize('x0', [0.0, 650.0, -650.0, 705.0, -720.0])
def test_exp_on_store_matches_numpy_exp(x0):
  # tqf_plan_paths(transform = EXP): every stored value within 4e-16 of exp(log-state)
  # (a hand-written table-driven exp was measured here and did not beat libdevice's in
  # the C5 generator, which is bound by the dispatch port: 3.27 ms against 3.15-3.19 ms)
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures, utils
  times = np.linspace(0.0, 1.0, 9)
  drift, vol = closures.affine_closures(0.1 - 0.5, 0.0, 1.0)
  spec = closures.resolve_spec(drift, vol)
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.05), dtype=np.float64)
  steps, record_slot = engine.record_plan(mask, 9)
  rng = engine.RngSpec(tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0)
  plan = engine.Plan(spec, all_times, steps, np.array([x0]), rng, 20000, np.float64)
  try:
    logs = plan.paths(record_slot, 9).cpu().numpy()
    got = plan.paths(record_slot, 9, exp_transform=True).cpu().numpy()
  finally:
    plan.close()
  with np.errstate(over='ignore', under='ignore'):
    want = np.exp(logs)
  assert np.abs(logs - x0).max() > 2.0            # the paths do spread over several units
  normal = want > 1e-300
  np.testing.assert_allclose(got[normal], want[normal], rtol=4e-16)
  np.testing.assert_allclose(got[~normal], want[~normal], rtol=1e-10, atol=1e-323)
