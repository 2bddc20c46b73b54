"""`tff.math.qmc`: digital nets, Sobol generating matrices,
lattice rules (`math/qmc/*.py`).

CPU: the oracle (`oracle/qmc.py`) against every known value the reference's
tests and docstrings hold (`sobol_test.py:64-190`, `digital_net_test.py:104-245`,
`lattice_rule_test.py:67-225`, `digital_net.py:58-68`), and libtqf's HOST table
builders (`tqf_qmc_sobol_generating_matrices`,
`tqf_qmc_scramble_generating_matrices`) against the oracle bit for bit.
GPU: the device samplers against the same known values and against the oracle
on seeded inputs - bit-exact (every operation is an integer XOR, one correctly
rounded cast and a division by a power of two; the lattice arithmetic is
correctly rounded operation by operation).
"""
import ctypes as C

import numpy as np
import pytest

from oracle import qmc as oqmc

# sobol_test.py:66-96 / digital_net_test.py:111-140: the first 29 points in 5 dimensions
SOBOL_29x5 = np.array([
    [0.00000, 0.00000, 0.00000, 0.00000, 0.00000], [0.50000, 0.50000, 0.50000, 0.50000, 0.50000],
    [0.25000, 0.75000, 0.75000, 0.75000, 0.25000], [0.75000, 0.25000, 0.25000, 0.25000, 0.75000],
    [0.12500, 0.62500, 0.37500, 0.12500, 0.12500], [0.62500, 0.12500, 0.87500, 0.62500, 0.62500],
    [0.37500, 0.37500, 0.62500, 0.87500, 0.37500], [0.87500, 0.87500, 0.12500, 0.37500, 0.87500],
    [0.06250, 0.93750, 0.56250, 0.31250, 0.68750], [0.56250, 0.43750, 0.06250, 0.81250, 0.18750],
    [0.31250, 0.18750, 0.31250, 0.56250, 0.93750], [0.81250, 0.68750, 0.81250, 0.06250, 0.43750],
    [0.18750, 0.31250, 0.93750, 0.43750, 0.56250], [0.68750, 0.81250, 0.43750, 0.93750, 0.06250],
    [0.43750, 0.56250, 0.18750, 0.68750, 0.81250], [0.93750, 0.06250, 0.68750, 0.18750, 0.31250],
    [0.03125, 0.53125, 0.90625, 0.96875, 0.96875], [0.53125, 0.03125, 0.40625, 0.46875, 0.46875],
    [0.28125, 0.28125, 0.15625, 0.21875, 0.71875], [0.78125, 0.78125, 0.65625, 0.71875, 0.21875],
    [0.15625, 0.15625, 0.53125, 0.84375, 0.84375], [0.65625, 0.65625, 0.03125, 0.34375, 0.34375],
    [0.40625, 0.90625, 0.28125, 0.09375, 0.59375], [0.90625, 0.40625, 0.78125, 0.59375, 0.09375],
    [0.09375, 0.46875, 0.46875, 0.65625, 0.28125], [0.59375, 0.96875, 0.96875, 0.15625, 0.78125],
    [0.34375, 0.71875, 0.71875, 0.40625, 0.03125], [0.84375, 0.21875, 0.21875, 0.90625, 0.53125],
    [0.21875, 0.84375, 0.09375, 0.53125, 0.40625]], dtype=np.float32)
INDICES = [1, 3, 10, 15, 19, 24, 28]                      # sobol_test.py:104-114
# sobol_test.py:127-135: 8 points in 6 dimensions, tent transform
SOBOL_TENT_8x6 = np.array([
    [0.00, 0.00, 0.00, 0.00, 0.00, 0.00], [1.00, 1.00, 1.00, 1.00, 1.00, 1.00],
    [0.50, 0.50, 0.50, 0.50, 0.50, 0.50], [0.50, 0.50, 0.50, 0.50, 0.50, 0.50],
    [0.25, 0.75, 0.75, 0.25, 0.25, 0.75], [0.75, 0.25, 0.25, 0.75, 0.75, 0.25],
    [0.75, 0.75, 0.75, 0.25, 0.75, 0.25], [0.25, 0.25, 0.25, 0.75, 0.25, 0.75]], dtype=np.float32)
# sobol_test.py:148-156
SOBOL_8x6 = np.array([
    [0.000, 0.000, 0.000, 0.000, 0.000, 0.000], [0.500, 0.500, 0.500, 0.500, 0.500, 0.500],
    [0.250, 0.750, 0.750, 0.750, 0.250, 0.250], [0.750, 0.250, 0.250, 0.250, 0.750, 0.750],
    [0.125, 0.625, 0.375, 0.125, 0.125, 0.375], [0.625, 0.125, 0.875, 0.625, 0.625, 0.875],
    [0.375, 0.375, 0.625, 0.875, 0.375, 0.125], [0.875, 0.875, 0.125, 0.375, 0.875, 0.625]])
# sobol_test.py:169-173
GENERATING_5x5 = np.array([[16, 8, 4, 2, 1], [16, 24, 20, 30, 17], [16, 24, 12, 18, 29],
                           [16, 24, 4, 10, 31], [16, 8, 4, 22, 31]])
# lattice_rule_test.py:28-32: n = 2^20 points in 20 dimensions
LATTICE_Z = [1, 387275, 314993, 50301, 174023, 354905, 303021, 486111, 286797, 463237,
             211171, 216757, 29831, 155061, 315509, 193933, 129563, 276501, 395079, 139111]
# lattice_rule_test.py:69-85
LATTICE_16x6 = np.array([
    [0.0000, 0.0000, 0.0000, 0.0000, 0.0000, 0.0000], [0.0625, 0.6875, 0.0625, 0.8125, 0.4375, 0.5625],
    [0.1250, 0.3750, 0.1250, 0.6250, 0.8750, 0.1250], [0.1875, 0.0625, 0.1875, 0.4375, 0.3125, 0.6875],
    [0.2500, 0.7500, 0.2500, 0.2500, 0.7500, 0.2500], [0.3125, 0.4375, 0.3125, 0.0625, 0.1875, 0.8125],
    [0.3750, 0.1250, 0.3750, 0.8750, 0.6250, 0.3750], [0.4375, 0.8125, 0.4375, 0.6875, 0.0625, 0.9375],
    [0.5000, 0.5000, 0.5000, 0.5000, 0.5000, 0.5000], [0.5625, 0.1875, 0.5625, 0.3125, 0.9375, 0.0625],
    [0.6250, 0.8750, 0.6250, 0.1250, 0.3750, 0.6250], [0.6875, 0.5625, 0.6875, 0.9375, 0.8125, 0.1875],
    [0.7500, 0.2500, 0.7500, 0.7500, 0.2500, 0.7500], [0.8125, 0.9375, 0.8125, 0.5625, 0.6875, 0.3125],
    [0.8750, 0.6250, 0.8750, 0.3750, 0.1250, 0.8750], [0.9375, 0.3125, 0.9375, 0.1875, 0.5625, 0.4375]],
                        dtype=np.float32)
LATTICE_INDICES = [2, 3, 6, 9, 11, 14]                    # lattice_rule_test.py:98-106
LATTICE_SHIFT = [.00, .05, .10, .15, .20, .25, .30, .35, .40, .45, .50, .55, .60, .65, .70, .75,
                 .80, .85, .90, .95]                       # lattice_rule_test.py:152-155
# lattice_rule_test.py:156-164
LATTICE_SHIFTED_8x5 = np.array([
    [0.000, 0.050, 0.100, 0.150, 0.200], [0.125, 0.425, 0.225, 0.775, 0.075],
    [0.250, 0.800, 0.350, 0.400, 0.950], [0.375, 0.175, 0.475, 0.025, 0.825],
    [0.500, 0.550, 0.600, 0.650, 0.700], [0.625, 0.925, 0.725, 0.275, 0.575],
    [0.750, 0.300, 0.850, 0.900, 0.450], [0.875, 0.675, 0.975, 0.525, 0.325]])
# lattice_rule_test.py:183-191
LATTICE_TENT_8x5 = np.array([
    [0.000, 0.000, 0.000, 0.000, 0.000], [0.250, 0.750, 0.250, 0.750, 0.250],
    [0.500, 0.500, 0.500, 0.500, 0.500], [0.750, 0.250, 0.750, 0.250, 0.750],
    [1.000, 1.000, 1.000, 1.000, 1.000], [0.750, 0.250, 0.750, 0.250, 0.750],
    [0.500, 0.500, 0.500, 0.500, 0.500], [0.250, 0.750, 0.250, 0.750, 0.250]], dtype=np.float32)


# --------------------------------------------------------------- CPU: oracle ----
def test_oracle_sobol_known_values():
  got = oqmc.sobol_sample(5, 29)
  assert got.dtype == np.float32
  np.testing.assert_allclose(got, SOBOL_29x5, rtol=1e-6)
  np.testing.assert_allclose(oqmc.sobol_sample(5, 29, sequence_indices=np.array(INDICES, np.int64)),
                             SOBOL_29x5[INDICES], rtol=1e-6)
  np.testing.assert_allclose(oqmc.sobol_sample(6, 8, apply_tent_transform=True), SOBOL_TENT_8x6,
                             rtol=1e-6)
  for dtype in (np.float32, np.float64):
    got = oqmc.sobol_sample(6, 8, dtype=dtype)
    assert got.dtype == dtype
    np.testing.assert_allclose(got, SOBOL_8x6, rtol=1e-6)


def test_oracle_generating_matrices_known_values():
  for dtype in (np.int32, np.int64):
    got = oqmc.sobol_generating_matrices(5, 31, 5, dtype=dtype)
    assert got.dtype == dtype
    np.testing.assert_array_equal(got, GENERATING_5x5)
  # digital_net_test.py:104-150: sampling the net from these matrices
  for dtype in (np.int32, np.int64):
    g = oqmc.sobol_generating_matrices(5, 29, 5, dtype=dtype)
    np.testing.assert_allclose(oqmc.digital_net_sample(g, 29, 5), SOBOL_29x5, rtol=1e-6)


def test_oracle_randomisation_known_values():
  # digital_net.py:58-68: the documented example pins TF's stateless integer uniform
  np.testing.assert_array_equal(oqmc.random_digital_shift(2, 10, (2, 3)), [586, 1011])
  # digital_net_test.py:27-101: ranges, shapes and dtypes
  for dtype in (np.int32, np.int64):
    shift = oqmc.random_digital_shift(6, 3, (2, 3), dtype=dtype)
    mats = oqmc.random_scrambling_matrices(6, 3, (2, 3), dtype=dtype)
    assert shift.shape == (6,) and mats.shape == (6, 3)
    assert shift.dtype == dtype and mats.dtype == dtype
    for a in (shift, mats):
      assert a.min() >= 4 and a.max() < 8
  # digital_net_test.py:268-295: scrambling with 2^(num_digits - 1) everywhere is a no-op
  for dtype in (np.int32, np.int64):
    g = oqmc.sobol_generating_matrices(6, 8, 3, dtype=dtype)
    same = oqmc.scramble_generating_matrices(g, np.full(g.shape, 4, dtype=dtype), 3, dtype=dtype)
    assert same.dtype == dtype
    np.testing.assert_array_equal(same, g)


def test_oracle_scrambled_net_keeps_net_property():
  # a linearly scrambled + shifted (0, m, s)-net in base 2 still has exactly one point in
  # every dyadic interval of length 2^-m of every coordinate
  m, dim = 7, 12
  s = oqmc.random_scrambling_matrices(dim, m, (7, 11))
  shift = oqmc.random_digital_shift(dim, m, (5, 3))
  pts = oqmc.sobol_sample(dim, 2**m, digital_shift=shift, scrambling_matrices=s, dtype=np.float64)
  cells = np.floor(pts * 2**m).astype(np.int64)
  for d in range(dim):
    assert sorted(cells[:, d]) == list(range(2**m))


def test_oracle_lattice_known_values():
  for dtype in (np.int32, np.int64):
    got = oqmc.lattice_rule_sample(np.array(LATTICE_Z, dtype=dtype), 6, 16)
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, LATTICE_16x6, rtol=1e-6)
  z = np.array(LATTICE_Z, dtype=np.int32)
  np.testing.assert_allclose(
      oqmc.lattice_rule_sample(z, 6, 16, sequence_indices=np.array(LATTICE_INDICES, np.int32)),
      LATTICE_16x6[LATTICE_INDICES], rtol=1e-6)
  for dtype in (np.float32, np.float64):
    got = oqmc.lattice_rule_sample(z, 5, 8, additive_shift=np.array(LATTICE_SHIFT, dtype=dtype),
                                   dtype=None if dtype == np.float32 else dtype)
    np.testing.assert_allclose(got, LATTICE_SHIFTED_8x5, rtol=1e-6, atol=1e-6)  # assertAllClose's default atol
    zero = oqmc.lattice_rule_sample(z, 5, 8, additive_shift=np.zeros(20, dtype=dtype), dtype=dtype)
    np.testing.assert_allclose(zero, oqmc.lattice_rule_sample(z, 5, 8, dtype=dtype), rtol=1e-6)
  np.testing.assert_allclose(oqmc.lattice_rule_sample(z, 5, 8, apply_tent_transform=True),
                             LATTICE_TENT_8x5, rtol=1e-6)
  v = oqmc.random_scrambling_vectors(20, (2, 3))
  assert v.shape == (20,) and v.dtype == np.float32 and v.min() >= 0 and v.max() < 1


def test_oracle_utils():
  # utils_test.py:27-105
  np.testing.assert_array_equal(oqmc.exp2(np.array([0, 1, 5, 30, 31, 40]), np.int32),
                                [1, 2, 32, 2**30, 2**31 - 1, 2**31 - 1])
  np.testing.assert_array_equal(oqmc.exp2(np.array([62, 63, 64]), np.int64),
                                [2**62, 2**63 - 1, 2**63 - 1])
  np.testing.assert_allclose(oqmc.log2(np.array([1., 2., 8., 1024.], np.float32)), [0, 1, 3, 10],
                             rtol=1e-6)
  np.testing.assert_allclose(oqmc.tent_transform(np.array([0., .25, .5, .75, 1.])),
                             [0., .5, 1., .5, 0.])
  np.testing.assert_array_equal(
      oqmc.filter_tensor(np.array([7, 7, 7, 7], np.int32), np.array([5, 5, 5, 5], np.int32),
                         np.array([0, 1, 2, 3], np.int32)), [7, 0, 7, 0])


# ------------------------------------------- CPU: libtqf host table builders ----
def _host_generating_matrices(dim, num_results, num_digits):
  from tff_b200.math.qmc import sobol as psobol
  return psobol.sobol_generating_matrices(dim, num_results, num_digits, dtype=np.int64)


@pytest.mark.parametrize('dim,num_results,num_digits', [
    (5, 31, 5), (1, 8, 3), (2, 2, 1), (40, 1000, 10), (300, 2**16 + 3, 20), (1111, 2**20, 31),
    (64, 2**24, 24), (21201, 64, 6)])
def test_host_generating_matrices_match_oracle(dim, num_results, num_digits):
  got = _host_generating_matrices(dim, num_results, num_digits)
  want = oqmc.sobol_generating_matrices(dim, num_results, num_digits, dtype=np.int64)
  assert got.shape == want.shape
  np.testing.assert_array_equal(got, want)


def test_host_generating_matrices_known_values_and_errors():
  from tff_b200.math import qmc
  for dtype in (np.int32, np.int64):
    got = qmc.sobol_generating_matrices(5, 31, 5, validate_args=True, dtype=dtype)
    assert got.dtype == dtype
    np.testing.assert_array_equal(got, GENERATING_5x5)
  with pytest.raises(ValueError):
    qmc.sobol_generating_matrices(0, 31, 5, validate_args=True)
  with pytest.raises(ValueError):
    qmc.sobol_generating_matrices(5, 0, 5, validate_args=True)
  with pytest.raises(ValueError):
    qmc.sobol_generating_matrices(5, 31, 0, validate_args=True)
  with pytest.raises(ValueError):
    qmc.sobol_generating_matrices(21202, 31, 5)


@pytest.mark.parametrize('int_dtype', [np.int32, np.int64])
def test_host_scrambling_matches_oracle(int_dtype):
  from tff_b200.math import qmc
  rng = np.random.default_rng(5)
  for dim, num_results, num_digits in [(6, 8, 3), (33, 5000, 13), (200, 2**20, 30)]:
    g = oqmc.sobol_generating_matrices(dim, num_results, num_digits, dtype=int_dtype)
    s = rng.integers(2**(num_digits - 1), 2**num_digits, size=(dim, num_digits)).astype(int_dtype)
    got = qmc.scramble_generating_matrices(g, s, num_digits, validate_args=True)
    assert got.dtype == int_dtype
    np.testing.assert_array_equal(got, oqmc.scramble_generating_matrices(g, s, num_digits))
    # digital_net_test.py:268-295
    same = qmc.scramble_generating_matrices(g, np.full(s.shape, 2**(num_digits - 1), int_dtype),
                                            num_digits, dtype=int_dtype)
    np.testing.assert_array_equal(same, g)


def test_qmc_utils_mirror():
  from tff_b200.math import qmc
  np.testing.assert_array_equal(qmc.utils.exp2(np.array([0, 1, 5, 30, 31, 40], np.int32)),
                                oqmc.exp2(np.array([0, 1, 5, 30, 31, 40]), np.int32))
  np.testing.assert_array_equal(qmc.utils.exp2(np.array([62, 63, 64], np.int64)),
                                oqmc.exp2(np.array([62, 63, 64]), np.int64))
  x = np.array([1., 2., 8., 1000.], np.float32)
  np.testing.assert_array_equal(qmc.utils.log2(x), oqmc.log2(x))
  np.testing.assert_array_equal(qmc.utils.tent_transform(np.array([0., .25, .5, .75, 1.])),
                                [0., .5, 1., .5, 0.])
  np.testing.assert_array_equal(
      qmc.utils.filter_tensor(np.array([7, 7, 7, 7], np.int32), np.array([5, 5, 5, 5], np.int32),
                              np.array([0, 1, 2, 3], np.int32)), [7, 0, 7, 0])
  for n in (1, 2, 3, 8, 29, 31, 1000, 2**20, 2**20 + 1):
    assert qmc.utils.ceil_log2_float32(n) == oqmc._ceil_log2_f32(n)  # pylint: disable=protected-access


# ---------------------------------------------------------------- GPU: device ----
def _np(t):
  return t.cpu().numpy()


@pytest.mark.gpu
def test_gpu_sobol_reference_kats():
  from tff_b200.math import qmc
  got = qmc.sobol_sample(5, 29, validate_args=True)
  assert _np(got).dtype == np.float32
  np.testing.assert_allclose(_np(got), SOBOL_29x5, rtol=1e-6)
  got = qmc.sobol_sample(5, 29, sequence_indices=np.array(INDICES, np.int64), validate_args=True)
  np.testing.assert_allclose(_np(got), SOBOL_29x5[INDICES], rtol=1e-6)
  np.testing.assert_allclose(_np(qmc.sobol_sample(6, 8, apply_tent_transform=True, validate_args=True)),
                             SOBOL_TENT_8x6, rtol=1e-6)
  for dtype in (np.float32, np.float64):
    got = _np(qmc.sobol_sample(6, 8, validate_args=True, dtype=dtype))
    assert got.dtype == dtype
    np.testing.assert_allclose(got, SOBOL_8x6, rtol=1e-6)
  # digital_net_test.py:104-245
  for int_dtype in (np.int32, np.int64):
    g = qmc.sobol_generating_matrices(5, 29, 5, dtype=int_dtype)
    np.testing.assert_allclose(_np(qmc.digital_net_sample(g, 29, 5, validate_args=True)), SOBOL_29x5,
                               rtol=1e-6)
  g = qmc.sobol_generating_matrices(5, 29, 5)
  got = qmc.digital_net_sample(g, 29, 5, sequence_indices=np.array(INDICES, np.int64), validate_args=True)
  np.testing.assert_allclose(_np(got), SOBOL_29x5[INDICES], rtol=1e-6)
  g = qmc.sobol_generating_matrices(6, 8, 3)
  np.testing.assert_allclose(_np(qmc.digital_net_sample(g, 8, 3, apply_tent_transform=True)),
                             SOBOL_TENT_8x6, rtol=1e-6)


@pytest.mark.gpu
def test_gpu_randomisation_matches_oracle():
  from tff_b200.math import qmc
  # digital_net.py:58-68
  np.testing.assert_array_equal(_np(qmc.random_digital_shift(2, 10, seed=(2, 3))), [586, 1011])
  for dtype in (np.int32, np.int64):
    for dim, nd, seed in [(6, 3, (2, 3)), (1, 1, (0, 0)), (257, 31, (123456789, -5)),
                          (1000, 20, (2**40, 7))]:
      if dtype == np.int32 and nd == 31:
        nd = 30
      got = _np(qmc.random_digital_shift(dim, nd, seed, dtype=dtype, validate_args=True))
      assert got.dtype == dtype and got.shape == (dim,)
      np.testing.assert_array_equal(got, oqmc.random_digital_shift(dim, nd, seed, dtype=dtype))
      assert got.min() >= 2**(nd - 1) and got.max() < 2**nd
      got = _np(qmc.random_scrambling_matrices(dim, nd, seed, dtype=dtype, validate_args=True))
      assert got.dtype == dtype and got.shape == (dim, nd)
      np.testing.assert_array_equal(got, oqmc.random_scrambling_matrices(dim, nd, seed, dtype=dtype))
  got = _np(qmc.random_digital_shift(300, 50, (9, 9), dtype=np.int64))
  np.testing.assert_array_equal(got, oqmc.random_digital_shift(300, 50, (9, 9), dtype=np.int64))
  with pytest.raises(ValueError):
    qmc.random_digital_shift(0, 3, (2, 3), validate_args=True)
  with pytest.raises(ValueError):
    qmc.random_digital_shift(3, 0, (2, 3), validate_args=True)
  for dtype in (np.float32, np.float64):
    v = _np(qmc.random_scrambling_vectors(20, (2, 3), dtype=dtype, validate_args=True))
    assert v.dtype == dtype and v.shape == (20,)
    np.testing.assert_array_equal(v, oqmc.random_scrambling_vectors(20, (2, 3), dtype=dtype))


@pytest.mark.gpu
@pytest.mark.parametrize('int_dtype', [np.int32, np.int64])
@pytest.mark.parametrize('real_dtype', [np.float32, np.float64])
def test_gpu_digital_net_matches_oracle(int_dtype, real_dtype):
  from tff_b200.math import qmc
  rng = np.random.default_rng(11)
  cases = [(5, 29, 5), (1, 1, 1), (3, 2, 1), (40, 1000, 10), (17, 4097, 13), (300, 2**16 + 3, 20),
           (64, 2**18, 30 if int_dtype == np.int32 else 40), (1200, 3000, 12)]
  for dim, num_results, num_digits in cases:
    g = qmc.sobol_generating_matrices(dim, num_results, num_digits, dtype=int_dtype)
    np.testing.assert_array_equal(g, oqmc.sobol_generating_matrices(dim, num_results, num_digits,
                                                                    dtype=int_dtype))
    shift = oqmc.random_digital_shift(dim, num_digits, (3, 4), dtype=int_dtype)
    scr = oqmc.random_scrambling_matrices(dim, num_digits, (5, 6), dtype=int_dtype)
    if g.shape[1] != num_digits:
      scr_for_sample = None              # the reference requires equal shapes
    else:
      scr_for_sample = scr
    for kw in (dict(), dict(digital_shift=shift), dict(apply_tent_transform=True),
               dict(digital_shift=shift, scrambling_matrices=scr_for_sample, apply_tent_transform=True)):
      got = _np(qmc.digital_net_sample(g, num_results, num_digits, dtype=real_dtype, **kw))
      want = oqmc.digital_net_sample(g, num_results, num_digits, dtype=real_dtype, **kw)
      assert got.dtype == real_dtype and got.shape == (num_results, dim)
      np.testing.assert_array_equal(got, want)
    if num_results > 4:
      idx = rng.integers(0, num_results, size=257)
      got = _np(qmc.digital_net_sample(g, num_results, num_digits, sequence_indices=idx,
                                       digital_shift=shift, dtype=real_dtype, validate_args=True))
      want = oqmc.digital_net_sample(g, num_results, num_digits, sequence_indices=idx,
                                     digital_shift=shift, dtype=real_dtype)
      np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_gpu_scrambled_sobol_matches_oracle_and_is_a_net():
  import torch
  from tff_b200.math import qmc
  m, dim = 12, 50
  n = 2**m
  shift = qmc.random_digital_shift(dim, m, (5, 3))
  scr = qmc.random_scrambling_matrices(dim, m, (7, 11))
  for dtype in (np.float32, np.float64):
    got = _np(qmc.sobol_sample(dim, n, digital_shift=shift, scrambling_matrices=scr, dtype=dtype,
                               validate_args=True))
    want = oqmc.sobol_sample(dim, n, digital_shift=_np(shift), scrambling_matrices=_np(scr), dtype=dtype)
    np.testing.assert_array_equal(got, want)
    cells = np.floor(got.astype(np.float64) * n).astype(np.int64)
    for d in range(dim):
      assert np.array_equal(np.sort(cells[:, d]), np.arange(n))
  # device-resident sequence indices are consumed in place
  idx = torch.arange(n - 1, -1, -1, device='cuda', dtype=torch.int64)
  got = _np(qmc.sobol_sample(dim, n, sequence_indices=idx, dtype=np.float64))
  np.testing.assert_array_equal(got, oqmc.sobol_sample(dim, n, dtype=np.float64)[::-1])
  # size-independent property at a size the oracle does not run: every coordinate of the
  # first 2^22 points is a permutation of the dyadic grid
  big = qmc.sobol_sample(8, 2**22, dtype=np.float64)
  cells = (big * 2**22).to(torch.int64)
  for d in range(8):
    assert bool((torch.sort(cells[:, d]).values == torch.arange(2**22, device='cuda')).all())


@pytest.mark.gpu
def test_gpu_digital_net_argument_errors():
  from tff_b200.math import qmc
  g = qmc.sobol_generating_matrices(5, 29, 5)
  with pytest.raises(ValueError):
    qmc.digital_net_sample(g[0], 29, 5, validate_args=True)
  with pytest.raises(ValueError):
    qmc.digital_net_sample(g, 0, 5, validate_args=True)
  with pytest.raises(ValueError):
    qmc.digital_net_sample(g, 29, 0, validate_args=True)
  with pytest.raises(ValueError):
    qmc.digital_net_sample(g, 29, 5, sequence_indices=[1, 29], validate_args=True)
  with pytest.raises(ValueError):
    qmc.digital_net_sample(g, 29, 5, digital_shift=[1, 2], validate_args=True)
  with pytest.raises(ValueError):
    qmc.digital_net_sample(g, 29, 5, scrambling_matrices=np.ones((5, 4), np.int32), validate_args=True)
  with pytest.raises(ValueError):
    qmc.digital_net_sample(g.astype(np.float32), 29, 5)


@pytest.mark.gpu
def test_gpu_lattice_rule_reference_kats_and_oracle():
  from tff_b200.math import qmc
  for int_dtype in (np.int32, np.int64):
    z = np.array(LATTICE_Z, dtype=int_dtype)
    got = _np(qmc.lattice_rule_sample(z, 6, 16, validate_args=True))
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, LATTICE_16x6, rtol=1e-6)
  z = np.array(LATTICE_Z, dtype=np.int32)
  got = qmc.lattice_rule_sample(z, 6, 16, sequence_indices=np.array(LATTICE_INDICES, np.int32),
                                validate_args=True)
  np.testing.assert_allclose(_np(got), LATTICE_16x6[LATTICE_INDICES], rtol=1e-6)
  for dtype in (np.float32, np.float64):
    got = _np(qmc.lattice_rule_sample(z, 5, 8, additive_shift=np.array(LATTICE_SHIFT, dtype=dtype),
                                      validate_args=True, dtype=dtype))
    assert got.dtype == dtype
    np.testing.assert_allclose(got, LATTICE_SHIFTED_8x5, rtol=1e-6, atol=1e-6)  # assertAllClose's default atol
    got = _np(qmc.lattice_rule_sample(z, 5, 8, additive_shift=np.zeros(20, dtype=dtype), dtype=dtype))
    np.testing.assert_allclose(got, LATTICE_SHIFTED_8x5 * 0 + _np(qmc.lattice_rule_sample(z, 5, 8, dtype=dtype)),
                               rtol=1e-6)
  np.testing.assert_allclose(_np(qmc.lattice_rule_sample(z, 5, 8, apply_tent_transform=True)),
                             LATTICE_TENT_8x5, rtol=1e-6)
  # bit-exact against the oracle, full 2^20-point rule in 20 dimensions, random shift
  rng = np.random.default_rng(3)
  for int_dtype in (np.int32, np.int64):
    for dtype in (np.float32, np.float64):
      zz = np.array(LATTICE_Z, dtype=int_dtype)
      shift = rng.random(20).astype(dtype) - (0.5 if dtype == np.float64 else 0.0)
      for kw in (dict(), dict(additive_shift=shift), dict(additive_shift=shift, apply_tent_transform=True)):
        got = _np(qmc.lattice_rule_sample(zz, 20, 2**20, dtype=dtype, **kw))
        want = oqmc.lattice_rule_sample(zz, 20, 2**20, dtype=dtype, **kw)
        np.testing.assert_array_equal(got, want)
      idx = rng.integers(0, 2**20, size=1000)
      got = _np(qmc.lattice_rule_sample(zz, 7, 2**20, sequence_indices=idx, additive_shift=shift, dtype=dtype))
      np.testing.assert_array_equal(
          got, oqmc.lattice_rule_sample(zz, 7, 2**20, sequence_indices=idx, additive_shift=shift, dtype=dtype))
  with pytest.raises(ValueError):
    qmc.lattice_rule_sample(z.reshape(4, 5), 5, 8, validate_args=True)
  with pytest.raises(ValueError):
    qmc.lattice_rule_sample(z, 21, 8, validate_args=True)
  with pytest.raises(ValueError):
    qmc.lattice_rule_sample(z, 5, 0, validate_args=True)


@pytest.mark.gpu
def test_gpu_qmc_normal_integral():
  # sobol_test.py:28-62: importance-sampled mean / stddev of N(mu_p, 0.5) under N(0, 1) draws
  import torch
  from tff_b200.math import qmc
  n = 1000
  u = qmc.sobol_sample(2, n + 1, sequence_indices=np.arange(1, n + 1), dtype=np.float64)
  q = torch.special.ndtri(u)
  mu_p = torch.tensor([-1., 1.], dtype=torch.float64, device='cuda')
  pdf = lambda x, mu, s: torch.exp(-0.5 * ((x - mu) / s)**2) / (s * np.sqrt(2 * np.pi))
  w = pdf(q, mu_p, 0.5) / pdf(q, 0.0, 1.0)
  e_x = (q * w).mean(0)
  std = torch.sqrt((q**2 * w - e_x**2).mean(0))
  np.testing.assert_allclose(_np(e_x), [-1., 1.], rtol=0.01)
  np.testing.assert_allclose(_np(std), [0.5, 0.5], rtol=0.02)
