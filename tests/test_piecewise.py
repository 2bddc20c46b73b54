"""`tff.math.piecewise` (host side): the reference's own tests, `math/piecewise_test.py:31-310`.

`PiecewiseConstantFunc` is the type model parameters arrive in on the hot path
(SURVEY 8a, row a17); it is evaluated on the host, once per grid time, into the
coefficient tables the kernels read.
"""
import numpy as np
import pytest

from tff_b200.math import piecewise


def test_find_interval_index():
  # piecewise_test.py:31-58
  assert isinstance(piecewise.find_interval_index([1.0], [0.0, 1.0])[0], np.int32)
  np.testing.assert_array_equal(piecewise.find_interval_index([1.0], [1.0]), [0])
  np.testing.assert_array_equal(piecewise.find_interval_index([0.0], [1.0]), [-1])
  np.testing.assert_array_equal(piecewise.find_interval_index([2.0], [1.0]), [0])
  np.testing.assert_array_equal(
      piecewise.find_interval_index([0.25, 3.0, 5.0, 0.0, 0.5, 0.8], [0.25, 0.5, 1.0, 2.0, 3.0]),
      [0, 4, 4, -1, 1, 1])
  np.testing.assert_array_equal(
      piecewise.find_interval_index([3.0, 4.0], [2.0, 3.0], last_interval_is_closed=True), [0, 1])


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_piecewise_constant_value_and_integral_no_batch(dtype):
  # piecewise_test.py:60-88
  f = piecewise.PiecewiseConstantFunc(np.array([0.1, 10], dtype=dtype), np.array([3, 4, 5], dtype=dtype), dtype=dtype)
  value = f(np.array([0., 0.1, 2., 11.]))
  assert value.dtype == dtype
  np.testing.assert_array_equal(value, [3., 3., 4., 5.])
  x = np.array([-4.1, 0., 1., 1.5, 2., 4.5, 5.5])
  f = piecewise.PiecewiseConstantFunc(np.array([1, 2, 3, 4, 5], dtype=dtype),
                                      np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6]), dtype=dtype)
  integral = f.integrate(x, x + 4.1)
  assert integral.dtype == dtype
  np.testing.assert_allclose(integral, [0.41, 1.05, 1.46, 1.66, 1.86, 2.41, 2.46], atol=1e-5, rtol=1e-5)


_X = np.array([[[0.0, 0.1, 2.0, 11.0], [0.0, 2.0, 3.0, 9.0]], [[0.0, 1.0, 2.0, 3.0], [4.0, 5.0, 6.0, 7.0]]])
_JUMPS = np.array([[[0.1, 10.0], [1.5, 10.0]], [[1.0, 2.0], [5.0, 6.0]]])
_VALUES = [[[3, 4, 5], [3, 4, 5]], [[3, 4, 5], [3, 4, 5]]]


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_piecewise_constant_value_and_integral_with_batch(dtype):
  # piecewise_test.py:90-113 (right-continuity) and :141-158
  f = piecewise.PiecewiseConstantFunc(_JUMPS, np.array(_VALUES, dtype=dtype), dtype=dtype)
  value = f(_X, left_continuous=False)
  assert value.dtype == dtype
  np.testing.assert_array_equal(value, [[[3.0, 4.0, 4.0, 5.0], [3.0, 4.0, 4.0, 4.0]],
                                        [[3.0, 4.0, 5.0, 5.0], [3.0, 4.0, 5.0, 5.0]]])
  integral = f.integrate(_X, _X + 1.1)
  assert integral.dtype == dtype
  np.testing.assert_allclose(integral, [[[4.3, 4.4, 4.4, 5.5], [3.3, 4.4, 4.4, 4.5]],
                                        [[3.4, 4.5, 5.5, 5.5], [3.4, 4.5, 5.5, 5.5]]], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_piecewise_constant_value_with_batch_and_repetitions(dtype):
  # piecewise_test.py:115-139
  x = np.array([[-4.1, 0.1, 1., 2., 10, 11.], [1., 2., 3., 2., 5., 9.]], dtype=dtype)
  jumps = np.array([[0.1, 0.1, 1., 1., 10., 10.], [-1., 1.2, 2.2, 2.2, 2.2, 8.]], dtype=dtype)
  values = np.array([[3, 3, 4, 5, 5., 2, 6.], [-1, -5, 2, 5, 5., 5., 1.]], dtype=dtype)
  f = piecewise.PiecewiseConstantFunc(jumps, values, dtype=dtype)
  value = f(x, left_continuous=True)
  assert value.dtype == dtype
  np.testing.assert_array_equal(value, [[3., 3., 4., 5., 5., 6.], [-5., 2., 5., 2., 5., 1.]])


@pytest.mark.parametrize('dtype', [np.float32, np.float64, None])
def test_invalid_shapes(dtype):
  # piecewise_test.py:160-181
  jumps = np.array([[0.1, 10], [2., 10]])
  with pytest.raises(ValueError):
    piecewise.PiecewiseConstantFunc(jumps, np.array([[[3, 4, 5], [3, 4, 5]]], dtype=dtype), dtype=dtype)
  with pytest.raises(ValueError):
    piecewise.PiecewiseConstantFunc(jumps, np.array([[3, 4, 5, 6], [3, 4, 5, 7]], dtype=dtype), dtype=dtype)


@pytest.mark.parametrize('dtype', [np.float32, np.float64, None])
def test_matrix_event_shape_no_batch_shape(dtype):
  # piecewise_test.py:183-211
  x = np.array([0., 0.1, 2., 11.])
  f = piecewise.PiecewiseConstantFunc([0.1, 10], [[[1, 2], [3, 4]], [[5, 6], [7, 8]], [[9, 10], [11, 12]]],
                                      dtype=dtype)
  assert f.dtype() == (np.float32 if dtype is None else dtype)
  np.testing.assert_allclose(f(x), [[[1, 2], [3, 4]], [[1, 2], [3, 4]], [[5, 6], [7, 8]], [[9, 10], [11, 12]]],
                             atol=1e-5, rtol=1e-5)
  np.testing.assert_allclose(f.integrate(x, x + 1),
                             [[[4.6, 5.6], [6.6, 7.6]], [[5, 6], [7, 8]], [[5, 6], [7, 8]], [[9, 10], [11, 12]]],
                             atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('dtype', [np.float32, np.float64, None])
def test_3d_event_shape_with_batch_shape(dtype):
  # piecewise_test.py:213-246 (and the dynamic-shape variant :248-285)
  x = np.array([[0, 1, 2, 3], [0.5, 1.5, 2.5, 3.5]])
  f = piecewise.PiecewiseConstantFunc([[0.5, 2], [0.5, 1.5]],
                                      [[[0, 1, 1.5], [2, 3, 0], [1, 0, 1]], [[0, 0.5, 1], [1, 3, 2], [2, 3, 1]]],
                                      dtype=dtype)
  np.testing.assert_allclose(f(x), [[[0, 1, 1.5], [2, 3, 0], [2, 3, 0], [1, 0, 1]],
                                    [[0, 0.5, 1], [1, 3, 2], [2, 3, 1], [2, 3, 1]]], atol=1e-5, rtol=1e-5)
  np.testing.assert_allclose(f.integrate(x, x + 1), [[[1, 2, 0.75], [2, 3, 0], [1, 0, 1], [1, 0, 1]],
                                                     [[1, 3, 2], [2, 3, 1], [2, 3, 1], [2, 3, 1]]],
                             atol=1e-5, rtol=1e-5)


def test_convert_to_tensor_or_func():
  # piecewise_test.py:287-310
  for i in [2.0, [1, 2, 3], np.arange(1, 5, 1)]:
    value, is_const = piecewise.convert_to_tensor_or_func(i, np.float64)
    assert isinstance(value, np.ndarray) and value.dtype == np.float64 and is_const
  pwc = piecewise.PiecewiseConstantFunc(np.arange(0, 10, 1), np.ones(11), dtype=np.float64)
  assert piecewise.convert_to_tensor_or_func(pwc) == (pwc, False)
