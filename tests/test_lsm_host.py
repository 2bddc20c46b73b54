"""Host side of the Longstaff-Schwartz mirror (CPU): the payoff and basis descriptors.

`make_basket_put_payoff` / `make_polynomial_basis` return descriptors the fused LSM
passes evaluate on the device; they stay callable on the host like the reference's
closures.  Here: the reference's `payoff_utils_test.py:48-97` on the oracle and on
the descriptors, and the descriptor basis (values and the exponent table the kernel
reads) against the oracle's restatement of `lsm.py:50-125`.
"""
import numpy as np
import pytest
import torch

from oracle import lsm as olsm
from tff_b200.models import longstaff_schwartz as lsm

# payoff_utils_test.py:26-33 (Longstaff & Schwartz 2001, table 1)
_SAMPLES = [[1.0, 1.09, 1.08, 1.34], [1.0, 1.16, 1.26, 1.54], [1.0, 1.22, 1.07, 1.03], [1.0, 0.93, 0.97, 0.92],
            [1.0, 1.11, 1.56, 1.52], [1.0, 0.76, 0.77, 0.90], [1.0, 0.92, 0.84, 1.01], [1.0, 0.88, 1.22, 1.34]]
_EXPECTED = [[0, 0], [0, 0], [0.07, 0.17], [0.18, 0.28], [0, 0], [0.2, 0.3], [0.09, 0.19], [0, 0]]


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_put_payoff_function(dtype):
  # payoff_utils_test.py:48-65
  paths = np.asarray(_SAMPLES, dtype=dtype)[..., None]
  np.testing.assert_allclose(olsm.make_basket_put_payoff([1.1, 1.2], dtype=dtype)(paths, 3), _EXPECTED,
                             rtol=1e-6, atol=1e-6)
  got = lsm.make_basket_put_payoff([1.1, 1.2], dtype=dtype)(torch.from_numpy(paths), 3)
  assert got.numpy().dtype == dtype
  np.testing.assert_allclose(got.numpy(), _EXPECTED, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_put_payoff_function_batch(dtype):
  # payoff_utils_test.py:75-97: one strike per batch element
  p1 = np.asarray(_SAMPLES, dtype=dtype)[..., None]
  paths = np.stack([p1, p1 + dtype(0.1)], axis=0)
  np.testing.assert_allclose(olsm.make_basket_put_payoff([1.1, 1.3], dtype=dtype)(paths, 3), _EXPECTED,
                             rtol=1e-6, atol=1e-6)
  got = lsm.make_basket_put_payoff([1.1, 1.3], dtype=dtype)(torch.from_numpy(paths), 3)
  np.testing.assert_allclose(got.numpy(), _EXPECTED, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('dim,degree', [(1, 2), (1, 3), (2, 2), (3, 2), (2, 10)])
@pytest.mark.parametrize('batch', [None, 3])
def test_polynomial_basis_descriptor_equals_the_oracle(dim, degree, batch):
  rs = np.random.RandomState(dim * 100 + degree)
  shape = (50, 4, dim) if batch is None else (batch, 50, 4, dim)
  paths = 1.0 + 0.2 * rs.standard_normal(shape)
  want = olsm.make_polynomial_basis(degree)(paths, 2)
  basis = lsm.make_polynomial_basis(degree)
  got = basis(torch.from_numpy(paths), 2).numpy()
  assert got.shape == want.shape == ((1 if batch is None else batch), (degree + 1)**dim, 50)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)
  # the exponent table handed to the kernel (`tqf_lsm_desc.exponents`): row k holds the powers of
  # basis function k, in the reference's `tf.meshgrid` ('xy') order
  e = basis.exponents(dim)
  assert e.shape == ((degree + 1)**dim, dim) and e.dtype == np.int32
  x = paths if batch is not None else paths[None]
  c = x[:, :, 2, :] - x[:, :, 2, :].mean(axis=1, keepdims=True)            # [B, N, dim]
  np.testing.assert_allclose(np.prod(c[:, None, :, :]**e[None, :, None, :], axis=-1), want, rtol=1e-12, atol=1e-14)


def test_polynomial_basis_reference_docstring_example():
  # lsm.py:62-77: degree 2 on two 2-d samples at time index 1
  paths = np.array([[[0.5, 0.3], [1.0, 1.0], [2.0, 1.5]], [[2.5, 1.2], [3.0, 2.0], [4.0, 1.8]]])   # [2, 3, 2]
  got = lsm.make_polynomial_basis(2)(torch.from_numpy(paths), 1).numpy()
  assert got.shape == (1, 9, 2)
  c = paths[:, 1, :] - paths[:, 1, :].mean(axis=0)
  np.testing.assert_allclose(got[0, 0], 1.0)
  # every basis function is a product of powers <= 2 of the centred coordinates
  table = {tuple(e): np.prod(c**e, axis=-1) for e in lsm.make_polynomial_basis(2).exponents(2)}
  assert len(table) == 9
  for k, e in enumerate(lsm.make_polynomial_basis(2).exponents(2)):
    np.testing.assert_allclose(got[0, k], table[tuple(e)], rtol=1e-14)
