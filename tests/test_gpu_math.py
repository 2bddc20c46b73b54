"""Accuracy of the hand-written FP64 device math (csrc/tqf_math.cuh) against
multiprecision references (mpmath), through the C-ABI test hook tqf_math_eval."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eval(fn, x):
  import torch
  from tff_b200 import _lib
  from tff_b200 import _tensor
  xin = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).cuda()
  out = torch.empty_like(xin)
  _lib.check(_lib.lib().tqf_math_eval(fn, xin.data_ptr(), out.data_ptr(),
                                      xin.numel(), _tensor.current_stream_ptr()))
  return out.cpu().numpy()


def _ulps(got, want):
  want = np.asarray(want, dtype=np.float64)
  return np.abs(got - want) / np.spacing(np.abs(want))


def _mp(fn, xs):
  import mpmath as mp
  mp.mp.dps = 40
  return np.array([float(fn(mp.mpf(float(x)))) for x in xs])


def test_log_pos():
  import mpmath as mp
  rs = np.random.RandomState(1)
  x = np.concatenate([rs.uniform(1e-7, 1, 3000), 1 - 10.0**rs.uniform(-16, -1, 2000),
                      2.0**rs.uniform(-60, 1, 2000), [1.0, 0.5, 2.0**-31, 1e-7]])
  got = _eval(0, x)
  want = _mp(mp.log, x)
  nz = want != 0
  assert _ulps(got[nz], want[nz]).max() <= 3.0
  assert got[~nz].tolist() == [0.0] * int((~nz).sum())


def test_sqrt_pos():
  rs = np.random.RandomState(2)
  x = np.concatenate([rs.uniform(0, 40, 4000), 10.0**rs.uniform(-290, 290, 4000)])
  got = _eval(1, x)
  assert _ulps(got, np.sqrt(x)).max() <= 1.0


def test_sincos_2pi():
  import mpmath as mp
  rs = np.random.RandomState(3)
  v = np.concatenate([rs.uniform(0, 2 * np.pi, 6000),
                      np.arange(9) * (np.pi / 4), [0.0, 2 * np.pi, 1e-300, 1e-9]])
  for fn, ref in ((3, mp.sin), (4, mp.cos)):
    got = _eval(fn, v)
    want = _mp(ref, v)
    # absolute accuracy relative to 1 ulp of 1.0 near the zeros, <= 2 ulp elsewhere
    err = np.abs(got - want)
    assert np.all((err <= 2.0 * np.spacing(np.abs(want))) | (err <= 2.3e-16))


def test_ndtri():
  import mpmath as mp
  rs = np.random.RandomState(4)
  u = np.concatenate([
      rs.uniform(0, 1, 6000),
      (rs.randint(1, 2**31, 4000).astype(np.float64)) / 2.0**31,   # Sobol-like
      2.0**-np.arange(1, 32), 1 - 2.0**-np.arange(2, 32),
      0.5 + np.array([0.0, 2.0**-32, -2.0**-32, 0.499, -0.499, 0.49903, -0.49904])])
  u = u[(u > 0) & (u < 1)]
  got = _eval(2, u)

  def ref(p):
    return mp.sqrt(2) * mp.erfinv(2 * p - 1)
  want = _mp(ref, u)
  nz = want != 0
  rel = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
  assert rel.max() <= 1e-15, rel.max()   # ~4 ulp (CUDA normcdfinv documents 5)
  assert np.all(got[~nz] == 0)
  # agreement with the oracle's scipy ndtri well inside the 1e-12 budget
  from scipy import special
  np.testing.assert_allclose(got, special.ndtri(u), rtol=2e-15, atol=0)
