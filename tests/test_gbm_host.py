"""GBM model classes, host side (CPU): `drift_fn()` / `volatility_fn()` of the mirrors of
`GeometricBrownianMotion` and `MultivariateGeometricBrownianMotion` evaluate on the host like
the reference's closures; the reference's own checks,
`geometric_brownian_motion_test.py:50-235`, on them and on the oracle's closures.
"""
import numpy as np
import pytest

import tff_b200 as tff
from oracle import models as omodels


def _np(x):
  return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_univariate_constant_drift_and_volatility(dtype):
  # geometric_brownian_motion_test.py:50-66
  process = tff.models.GeometricBrownianMotion(0.05, 0.5, dtype=dtype)
  state = np.array([[1.], [2.], [3.]], dtype=dtype)
  for drift_fn, vol_fn in ((process.drift_fn(), process.volatility_fn()), omodels.gbm_closures(0.05, 0.5, dtype)):
    np.testing.assert_allclose(_np(drift_fn(0.2, state)), state * 0.05, atol=1e-8, rtol=1e-8)
    np.testing.assert_allclose(_np(vol_fn(0.2, state)), (state * 0.5)[..., None], atol=1e-8, rtol=1e-8)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_univariate_time_varying_drift_and_volatility(dtype):
  # geometric_brownian_motion_test.py:87-113 (one time per call: the samplers pass a scalar t)
  times = np.linspace(0.0, 10.0, 6, dtype=dtype)
  drift = np.append([0.0], np.sin(times, dtype=dtype)).astype(dtype)
  sigma = (np.append([0.0], np.cos(times, dtype=dtype)) ** 2.0).astype(dtype)
  pw = tff.math.piecewise.PiecewiseConstantFunc
  drift_in, sigma_in = pw(times, drift, dtype=dtype), pw(times, sigma, dtype=dtype)
  process = tff.models.GeometricBrownianMotion(drift_in, sigma_in, dtype=dtype)
  odrift, ovol = omodels.gbm_closures(omodels.PiecewiseConstantFunc(times, drift, dtype=dtype),
                                      omodels.PiecewiseConstantFunc(times, sigma, dtype=dtype), dtype)
  state = np.array([[1.], [2.], [3.]], dtype=dtype)
  for t in np.array([1.0, 3.5, 7.5, 9.8, 12], dtype=dtype):
    for drift_fn, vol_fn in ((process.drift_fn(), process.volatility_fn()), (odrift, ovol)):
      np.testing.assert_allclose(_np(drift_fn(t, state)), drift_in(t) * state, atol=1e-8, rtol=1e-8)
      np.testing.assert_allclose(_np(vol_fn(t, state)), sigma_in(t) * state[..., None], atol=1e-8, rtol=1e-8)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('corr_matrix', [[[1, 0.1], [0.1, 1]], None])
def test_multivariate_drift_and_volatility(dtype, corr_matrix):
  # geometric_brownian_motion_test.py:182-235
  means, vols = [0.05, 0.02], [0.1, 0.2]
  process = tff.models.MultivariateGeometricBrownianMotion(
      dim=2, means=means, volatilities=vols, corr_matrix=corr_matrix, dtype=np.float64)
  state = np.array([[1., 2.], [3., 4.], [5., 6.]], dtype=dtype)
  chol = np.linalg.cholesky(np.eye(2) if corr_matrix is None else corr_matrix)
  expected_vol = (np.array(vols) * state)[..., None] * chol
  odrift, ovol = omodels.mvgbm_closures(np.array(means), np.array(vols), corr_matrix, np.float64)
  # (a float32 state stays float32 here; TensorFlow promotes it to the process dtype)
  tol = 1e-8 if dtype == np.float64 else 1e-6
  for drift_fn, vol_fn in ((process.drift_fn(), process.volatility_fn()), (odrift, ovol)):
    np.testing.assert_allclose(_np(drift_fn(0.2, state)), np.array(means) * state, atol=tol, rtol=tol)
    np.testing.assert_allclose(_np(vol_fn(0.2, state)), expected_vol, atol=tol, rtol=tol)


def test_closures_carry_the_device_model_spec():
  # what `euler_sampling.sample` recognises instead of running Python inside a kernel
  from tff_b200.models import closures
  gbm = tff.models.GeometricBrownianMotion(0.05, 0.5, dtype=np.float64)
  spec = closures.resolve_spec(gbm.drift_fn(), gbm.volatility_fn(), dim=1)
  assert spec is gbm.drift_fn().tqf_spec
  mv = tff.models.MultivariateGeometricBrownianMotion(dim=2, means=[0.05, 0.02], volatilities=[0.1, 0.2],
                                                      corr_matrix=[[1, 0.1], [0.1, 1]], dtype=np.float64)
  assert closures.resolve_spec(mv.drift_fn(), mv.volatility_fn(), dim=2) is mv.drift_fn().tqf_spec
  with pytest.raises(NotImplementedError):
    closures.resolve_spec(gbm.drift_fn(), lambda t, x: x, dim=1)
