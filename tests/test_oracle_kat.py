"""Pins the CPU oracle against the reference's own known-answer tests.

Each test cites the reference test (file:line under
/root/reference/tf_quant_finance) that holds the expected values.
"""
import os

import numpy as np
import pytest

from oracle import draws
from oracle import grid
from oracle import philox
from oracle import sobol

REF_SOBOL = '/root/reference/third_party/sobol_data/new-joe-kuo-6.21201'


# ---------------------------------------------------------------- Sobol ----
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_sobol_known_values_small_dimension(dtype):
  # math/random_ops/sobol/sobol_test.py:28-38
  expected = np.array([[0.5, 0.5], [0.25, 0.75], [0.75, 0.25],
                       [0.125, 0.625], [0.625, 0.125]], dtype=dtype)
  got = sobol.sample(2, 5, dtype=dtype)
  assert got.dtype == dtype
  np.testing.assert_array_equal(got, expected)


def test_sobol_more_known_values():
  # math/random_ops/sobol/sobol_test.py:40-81 (compared as a set of rows)
  expected = [[0.5, 0.5, 0.5, 0.5, 0.5], [0.75, 0.25, 0.25, 0.25, 0.75],
              [0.25, 0.75, 0.75, 0.75, 0.25],
              [0.375, 0.375, 0.625, 0.875, 0.375],
              [0.875, 0.875, 0.125, 0.375, 0.875],
              [0.625, 0.125, 0.875, 0.625, 0.625],
              [0.125, 0.625, 0.375, 0.125, 0.125],
              [0.1875, 0.3125, 0.9375, 0.4375, 0.5625],
              [0.6875, 0.8125, 0.4375, 0.9375, 0.0625],
              [0.9375, 0.0625, 0.6875, 0.1875, 0.3125],
              [0.4375, 0.5625, 0.1875, 0.6875, 0.8125],
              [0.3125, 0.1875, 0.3125, 0.5625, 0.9375],
              [0.8125, 0.6875, 0.8125, 0.0625, 0.4375],
              [0.5625, 0.4375, 0.0625, 0.8125, 0.1875],
              [0.0625, 0.9375, 0.5625, 0.3125, 0.6875],
              [0.09375, 0.46875, 0.46875, 0.65625, 0.28125],
              [0.59375, 0.96875, 0.96875, 0.15625, 0.78125],
              [0.84375, 0.21875, 0.21875, 0.90625, 0.53125],
              [0.34375, 0.71875, 0.71875, 0.40625, 0.03125],
              [0.46875, 0.09375, 0.84375, 0.28125, 0.15625],
              [0.96875, 0.59375, 0.34375, 0.78125, 0.65625],
              [0.71875, 0.34375, 0.59375, 0.03125, 0.90625],
              [0.21875, 0.84375, 0.09375, 0.53125, 0.40625],
              [0.15625, 0.15625, 0.53125, 0.84375, 0.84375],
              [0.65625, 0.65625, 0.03125, 0.34375, 0.34375],
              [0.90625, 0.40625, 0.78125, 0.59375, 0.09375],
              [0.40625, 0.90625, 0.28125, 0.09375, 0.59375],
              [0.28125, 0.28125, 0.15625, 0.21875, 0.71875],
              [0.78125, 0.78125, 0.65625, 0.71875, 0.21875],
              [0.53125, 0.03125, 0.40625, 0.46875, 0.46875],
              [0.03125, 0.53125, 0.90625, 0.96875, 0.96875]]
  got = sobol.sample(5, 31, dtype=np.float32)
  assert sorted(map(tuple, expected)) == sorted(map(tuple, got.tolist()))


def test_sobol_skip():
  # math/random_ops/sobol/sobol_test.py:83-91
  a = sobol.sample(10, 67, dtype=np.float32)
  b = sobol.sample(10, 50, skip=17, dtype=np.float32)
  np.testing.assert_array_equal(a[17:], b)


def test_sobol_large_skip():
  # math/random_ops/sobol/sobol_test.py:93-98
  got = sobol.sample(1, 3, skip=2**31 - 5, dtype=np.float32)
  np.testing.assert_array_equal(got, [[0.25], [0.75], [0.5]])


@pytest.mark.skipif(not os.path.exists(REF_SOBOL),
                    reason='reference checkout absent (GPU box)')
def test_packed_joe_kuo_equals_reference_text():
  poly_t, m_t = sobol.parse_joe_kuo_text(REF_SOBOL)
  poly, m = sobol.load_joe_kuo()
  np.testing.assert_array_equal(poly, poly_t)
  np.testing.assert_array_equal(m, m_t)


# ------------------------------------------------- Sobol -> normal layout ----
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_sobol_numbers_generation(dtype):
  # models/utils_test.py:33-55
  samples = draws.generate_mc_normal_draws(
      num_normal_draws=2, num_time_steps=3, num_sample_paths=4,
      random_type=draws.RandomType.SOBOL, dtype=dtype, skip=10)
  expected = [[[0.8871465, 0.48877636], [-0.8871465, -0.48877636],
               [0.48877636, 0.8871465], [-0.15731068, 0.15731068]],
              [[0.8871465, -1.5341204], [1.5341204, -0.15731068],
               [-0.15731068, 1.5341204], [-0.8871465, 0.48877636]],
              [[-0.15731068, 1.5341204], [0.15731068, -0.48877636],
               [-1.5341204, 0.8871465], [0.8871465, -1.5341204]]]
  assert samples.dtype == dtype
  np.testing.assert_allclose(samples, expected, rtol=1e-5, atol=1e-5)
  # the flat index rule of SURVEY a3: (p, s, j) -> dimension s*dim + j of
  # point skip+1+p
  u = sobol.sample(6, 4, skip=10, dtype=np.float64)
  from scipy import special
  z = special.ndtri(u).reshape(4, 3, 2).transpose(1, 0, 2)
  np.testing.assert_allclose(samples, z, rtol=2e-6 if dtype == np.float32 else 1e-14)


# ----------------------------------------------------------------- grid ----
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_prepare_grid_num_time_step(dtype):
  # models/utils_test.py:102-116
  times = grid.tf_linspace(0.02, 1.0, 50, dtype)
  time_step = times[-1] / dtype(100)
  g, _, idx = grid.prepare_grid(times=times, time_step=time_step, dtype=dtype,
                                num_time_steps=100)
  np.testing.assert_allclose(g, np.linspace(0, 1, 101, dtype=dtype),
                             rtol=1e-6, atol=1e-6)
  np.testing.assert_array_equal(idx, [2 * i for i in range(1, 51)])


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_prepare_grid_time_step(dtype):
  # models/utils_test.py:122-132
  times = np.array([0.1, 0.5, 1, 2], dtype=dtype)
  g, mask, idx = grid.prepare_grid(times=times, time_step=0.1, dtype=dtype)
  np.testing.assert_allclose(g, np.linspace(0, 2, 21, dtype=dtype),
                             rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(g[idx], times, rtol=1e-6, atol=1e-6)
  assert mask.sum() == 4 and not mask[0]


def test_grid_sizes_of_the_configs():
  # S = 100 (C1), 252 (C2), 360 (C3)
  g, m, _ = grid.euler_grid([1.0], dtype=np.float64, time_step=0.01)
  assert g.shape[0] == 101 and m[-1] and m.sum() == 1
  g, m, _ = grid.euler_grid([1.0], dtype=np.float64, num_time_steps=252)
  assert g.shape[0] == 253 and g[0] == 0 and g[-1] == 1.0
  g, m, _ = grid.euler_grid([1.0], dtype=np.float32, num_time_steps=252)
  assert g.shape[0] == 253


# --------------------------------------------------------------- Philox ----
def test_philox_core_random123_kat():
  # Random123 kat_vectors, philox4x32 10 rounds.
  def run(ctr, key):
    return [int(x) for x in philox.philox4x32_10(
        np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))]
  assert run([0, 0, 0, 0], [0, 0]) == [
      0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
  assert run([0xffffffff] * 4, [0xffffffff] * 2) == [
      0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
  assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
             [0xa4093822, 0x299f31d0]) == [
                 0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_counter_add_carries():
  c = np.array([0xfffffffe, 0xffffffff, 0x7, 0x1], dtype=np.uint32)
  got = philox.counter_add(c, np.array([0, 1, 2, 3], dtype=np.uint64))
  assert got.tolist() == [[0xfffffffe, 0xffffffff, 7, 1],
                          [0xffffffff, 0xffffffff, 7, 1],
                          [0, 0, 8, 1], [1, 0, 8, 1]]


# Values TensorFlow itself publishes for these streams (TensorFlow "Random number
# generation" guide and the API docs of tf.random.stateless_normal): the seed
# scrambling (GenerateKey), the group / counter layout, Uint32ToFloat and
# BoxMullerFloat of oracle/philox.py are pinned by them.
TF_STATELESS_NORMAL_2x3_SEED_1_2 = np.array(
    [[0.5441101, 0.20738031, 0.07356433], [0.04643455, -1.3015898, -0.95385665]], np.float32)
TF_GENERATOR_FROM_SEED_1_NORMAL_2x3 = np.array(
    [[0.43842277, -0.53439844, -0.07710262], [1.5658046, -0.1012345, -0.2744976]], np.float32)


def test_philox_tensorflow_published_stateless_normal():
  # tf.random.stateless_normal([2, 3], seed=[1, 2]) as printed in TensorFlow's docs
  got = philox.stateless_normal([2, 3], [1, 2], np.float32)
  np.testing.assert_allclose(got, TF_STATELESS_NORMAL_2x3_SEED_1_2, rtol=2e-7, atol=1e-8)


def test_philox_tensorflow_published_generator_normal():
  # tf.random.Generator.from_seed(1).normal([2, 3]) (RNG guide): state = [1, 0, 0]
  # -> counter = [1, 0, 0, 0], key = [0, 0]; same Philox / Box-Muller kernels
  got = philox.normal_fill(np.array([0, 0], np.uint32), np.array([1, 0, 0, 0], np.uint32),
                           6, np.float32).reshape(2, 3)
  np.testing.assert_allclose(got, TF_GENERATOR_FROM_SEED_1_NORMAL_2x3, rtol=2e-7, atol=1e-8)


def test_philox_fp64_stream_is_consistent_with_the_pinned_fp32_one():
  # No TensorFlow publication holds float64 values.  What can be pinned without
  # TF: (i) float64 consumes the SAME raw words, two groups of two words per pair
  # of normals; (ii) Uint64ToDouble is the bit layout of random_distributions.h
  # (mantissa = low 20 bits of word 0 | word 1); (iii) BoxMullerDouble inverts:
  # z0^2 + z1^2 = -2 ln u1 and atan2(z0, z1) = 2 pi u2 for the uniforms of (ii).
  key, ctr = philox.stateless_key_counter([1, 2])
  words = philox.raw_words(key, ctr, 0, 4096)
  z = philox.normal_fill(key, ctr, 8192, np.float64).reshape(-1, 2)
  u1 = philox.uint64_to_double(words[:, 0], words[:, 1])
  u2 = philox.uint64_to_double(words[:, 2], words[:, 3])
  bits = (u1 + 1.0).view(np.uint64)
  np.testing.assert_array_equal(bits >> np.uint64(52), 1023)
  np.testing.assert_array_equal(
      bits & np.uint64((1 << 52) - 1),
      ((words[:, 0].astype(np.uint64) & np.uint64(0xFFFFF)) << np.uint64(32))
      | words[:, 1].astype(np.uint64))
  np.testing.assert_allclose((z**2).sum(axis=1), -2 * np.log(np.maximum(u1, 1e-7)), rtol=1e-13)
  ang = np.arctan2(z[:, 0], z[:, 1]) % (2 * np.pi)
  np.testing.assert_allclose(ang, 2 * np.pi * u2, rtol=0, atol=1e-12)
  # the float32 stream of the same seed starts from the same first group
  f = philox.normals_from_words(words[:1], np.float32)
  np.testing.assert_allclose(f[:3], TF_STATELESS_NORMAL_2x3_SEED_1_2[0], rtol=2e-7)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_stateless_prefix_stability_and_moments(dtype):
  # math/random_ops/multivariate_normal_test.py:342-380 (structure only)
  a = philox.stateless_normal([1000, 6], [4, 2], dtype)
  b = philox.stateless_normal([2000, 6], [4, 2], dtype)
  np.testing.assert_array_equal(a, b[:1000])
  big = philox.stateless_normal([200000], [1, 7], dtype).astype(np.float64)
  assert abs(big.mean()) < 1e-2 and abs(big.std() - 1) < 1e-2


def test_antithetic_pairing():
  # math/random_ops/multivariate_normal_test.py:249-281
  mean = np.zeros(6)
  z = draws.mv_normal_sample([10], mean,
                             random_type=draws.RandomType.STATELESS_ANTITHETIC,
                             seed=[1, 2], dtype=np.float64)
  np.testing.assert_allclose(z[:5] + z[5:], 0.0, atol=1e-10)
  d = draws.generate_mc_normal_draws(
      2, 3, 10, draws.RandomType.STATELESS_ANTITHETIC, seed=[1, 2],
      dtype=np.float64)
  assert d.shape == (3, 10, 2)
  np.testing.assert_array_equal(d[:, :5], -d[:, 5:])
  np.testing.assert_array_equal(d[1, 3], z[3].reshape(3, 2)[1])


# ----------------------------------------------------------- Hull-White ----
def _flat_rate(t):
  return 0.01 + 0 * t


def test_hw_discount_bond_price_kat():
  # models/hull_white/hull_white_test.py:465-485 (atol 1e-12 there)
  from oracle import hull_white
  m = hull_white.HullWhiteModel1F(0.1, 0.01, _flat_rate, np.float64)
  got = m.discount_bond_price([[0.011], [0.01]], [1.0, 2.0], [2.0, 3.5])
  np.testing.assert_allclose(got[:, 0], [0.98906753, 0.98495442], atol=5e-9)


def test_hw_exact_moments():
  # models/hull_white/hull_white_test.py:42-58, 100-131 (Brigo-Mercurio)
  from oracle import hull_white
  from oracle import models as omodels
  a, sigma = 0.1, 0.01
  vol = omodels.PiecewiseConstantFunc([0.1, 2.0], 3 * [sigma], dtype=np.float64)
  m = hull_white.HullWhiteModel1F(a, vol, _flat_rate, np.float64)
  paths = m.sample_paths([0.1, 0.5, 1.0], 50000, draws.RandomType.SOBOL,
                         skip=1000000)
  assert paths.shape == (50000, 3, 1)
  x = paths[:, -1, 0]
  true_mean = (0.01 + (sigma**2 / 2 / a**2) * (1 - np.exp(-a))**2)
  true_var = sigma**2 / 2 / a * (1 - np.exp(-2 * a))
  np.testing.assert_allclose(x.mean(), true_mean, rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(x.var(), true_var, rtol=1e-4, atol=1e-4)


def test_hw_swaption_mc_kat():
  # models/hull_white/swaption_test.py:85-125: analytic 0.71632434, the
  # reference's own MC tolerance is 1e-3 with 500k STATELESS_ANTITHETIC paths.
  from oracle import hull_white
  price = hull_white.swaption_price_mc(
      expiries=np.array(1.0),
      fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
      fixed_leg_daycount_fractions=0.25 * np.ones(4),
      fixed_leg_coupon=0.011 * np.ones(4), reference_rate_fn=_flat_rate,
      notional=100., mean_reversion=0.03, volatility=0.02, num_samples=500000,
      time_step=0.1, random_type=draws.RandomType.STATELESS_ANTITHETIC,
      seed=[4, 2], dtype=np.float64)
  assert price.shape == ()
  np.testing.assert_allclose(price, 0.71632434, rtol=1e-3, atol=1e-3)




def test_vector_hw_2d_moments_kat():
  # models/hull_white/hull_white_test.py:223-270: two correlated factors with
  # (batched) piecewise-constant volatilities; terminal mean / variance against
  # the Brigo-Mercurio closed forms (1e-4) and the correlation (1e-2).
  from oracle import hull_white
  from oracle import models
  a = np.array([0.1, 0.05]); sigma = np.array([0.01, 0.02])
  vols = [models.PiecewiseConstantFunc([0.1, 0.2, 0.5], 4 * [sigma[0]]),
          models.PiecewiseConstantFunc([0.1, 2.0, 3.0], 4 * [sigma[1]])]
  m = hull_white.VectorHullWhiteModel(2, a, vols, _flat_rate, [[1., 0.5], [0.5, 1.]])
  paths = m.sample_paths([0.1, 0.5, 1.0], 50000, draws.RandomType.STATELESS_ANTITHETIC,
                         seed=[1, 2])
  assert paths.shape == (50000, 3, 2)
  x = paths[:, -1, :]
  true_mean = 0.01 + sigma**2 / 2 / a**2 * (1 - np.exp(-a))**2
  true_var = sigma**2 / 2 / a * (1 - np.exp(-2 * a))
  np.testing.assert_allclose(x.mean(0), true_mean, rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(x.var(0), true_var, rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(np.corrcoef(x[:, 0], x[:, 1])[0, 1], 0.5, atol=1e-2)


def test_hw_bond_option_mc_kat():
  # models/hull_white/zero_coupon_bond_option_test.py:49-73 (analytic 0.02817777;
  # the reference's MC tolerance is 1e-4 with 500k PSEUDO_ANTITHETIC paths)
  from oracle import hull_white
  expiries, maturities = np.array(1.0), np.array(5.0)
  strikes = np.exp(-0.01 * maturities) / np.exp(-0.01 * expiries)
  price = hull_white.bond_option_price_mc(
      strikes=strikes, expiries=expiries, maturities=maturities,
      discount_rate_fn=_flat_rate, mean_reversion=0.03, volatility=0.02,
      num_samples=500000, time_step=0.1,
      random_type=draws.RandomType.STATELESS_ANTITHETIC, seed=[1, 7])
  assert price.shape == ()
  np.testing.assert_allclose(price, 0.02817777, rtol=1e-4, atol=1e-4)


def test_hw_cap_mc_kat():
  # models/hull_white/cap_floor_test.py:57-83: 0.4072088281493774 +- 1e-3 with
  # 50k STATELESS_ANTITHETIC paths, seed [42, 42]; first caplet expires at t = 0.
  from oracle import hull_white
  price = hull_white.cap_floor_price_mc(
      strikes=0.01 * np.ones(4), expiries=np.array([0.0, 0.25, 0.5, 0.75]),
      maturities=np.array([0.25, 0.5, 0.75, 1.0]),
      daycount_fractions=0.25 * np.ones(4), notional=100.0,
      reference_rate_fn=_flat_rate, mean_reversion=0.03, volatility=0.02,
      num_samples=50_000, time_step=0.1,
      random_type=draws.RandomType.STATELESS_ANTITHETIC, seed=[42, 42])
  assert price.shape == ()
  np.testing.assert_allclose(price, 0.4072088281493774, rtol=1e-3, atol=1e-3)



def _bermudan_legs(exercise):
  """`bermudan_swaption_test.py:31-80`: semi-annual 5y swap, payments clipped at 5y."""
  start = np.array([np.clip(np.arange(8) * 0.5 + e, None, 5.0) for e in exercise])
  end = np.clip(start + 0.5, 0.0, 5.0)
  return dict(exercise_times=np.array(exercise), fixed_leg_payment_times=end,
              fixed_leg_daycount_fractions=end - start,
              fixed_leg_coupon=0.011 * np.ones_like(end))


def test_hw_bermudan_swaption_mc_kat():
  # models/hull_white/bermudan_swaption_test.py:86-183: 5nc1 1.8892 (tol 1e-2,
  # 10k paths) and the [5nc1, 5nc2] batch [1.8892, 1.6633] (tol 5e-3, 50k paths),
  # STATELESS_ANTITHETIC seed [0, 0], time_step 0.1
  from oracle import hull_white
  ex1 = [1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5]
  ex2 = [2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 5.0]
  kw = dict(reference_rate_fn=_flat_rate, notional=100., mean_reversion=0.03,
            volatility=0.01, time_step=0.1,
            random_type=draws.RandomType.STATELESS_ANTITHETIC, seed=[0, 0])
  price = hull_white.bermudan_swaption_price_mc(num_samples=10000, **_bermudan_legs(ex1), **kw)
  assert price.shape == ()
  np.testing.assert_allclose(price, 1.8892, rtol=1e-2, atol=1e-2)
  l1, l2 = _bermudan_legs(ex1), _bermudan_legs(ex2)
  legs = {k: np.stack([l1[k], l2[k]]) for k in l1}
  price = hull_white.bermudan_swaption_price_mc(num_samples=50000, **legs, **kw)
  assert price.shape == (2,)
  np.testing.assert_allclose(price, [1.8892, 1.6633], rtol=5e-3, atol=5e-3)


# ------------------------------------------------------ Longstaff-Schwartz ----
_LS_SAMPLES = np.expand_dims([[1.0, 1.09, 1.08, 1.34], [1.0, 1.16, 1.26, 1.54],
                              [1.0, 1.22, 1.07, 1.03], [1.0, 0.93, 0.97, 0.92],
                              [1.0, 1.11, 1.56, 1.52], [1.0, 0.76, 0.77, 0.90],
                              [1.0, 0.92, 0.84, 1.01], [1.0, 0.88, 1.22, 1.34]], -1)
_LS_DF = np.exp(-np.cumsum([0.06, 0.06, 0.06]))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_lsm_reference_kats(dtype):
  # models/longstaff_schwartz/lsm_test.py:60-127, 169-219
  from oracle import lsm
  basis = lsm.make_polynomial_basis(2)
  put = lsm.make_basket_put_payoff([1.1], dtype=dtype)
  tol = dict(rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(
      lsm.least_square_mc(_LS_SAMPLES, [3], put, basis, [_LS_DF[-1]], dtype=dtype),
      [0.0564], **tol)
  np.testing.assert_allclose(
      lsm.least_square_mc(_LS_SAMPLES, [1, 2, 3], put, basis, _LS_DF, dtype=dtype),
      [0.1144], **tol)
  np.testing.assert_allclose(
      lsm.least_square_mc(_LS_SAMPLES, [1, 2, 3], put, basis, _LS_DF,
                          num_calibration_samples=4, dtype=dtype), [0.174226], **tol)
  put2 = lsm.make_basket_put_payoff([1.1, 1.2], dtype=dtype)
  df2 = np.exp(-np.cumsum([[0.06] * 3, [0.05] * 3], -1))[None]
  np.testing.assert_allclose(
      lsm.least_square_mc(_LS_SAMPLES, [1, 2, 3], put2, basis, df2, dtype=dtype),
      [0.1144, 0.199], **tol)
  batch = np.stack([_LS_SAMPLES, _LS_SAMPLES + 0.1], 0)
  np.testing.assert_allclose(
      lsm.least_square_mc(batch, [1, 2, 3], put2, basis, df2, dtype=dtype),
      [0.1144, 0.1157], **tol)


def test_lsm_basket_degree_10():
  # models/longstaff_schwartz/lsm_test.py:129-157 (rank-deficient regression)
  from oracle import lsm
  basis = lsm.make_polynomial_basis(10)
  put = lsm.make_basket_put_payoff([1.1, 1.2, 1.3], dtype=np.float64)
  s2 = np.concatenate([_LS_SAMPLES, _LS_SAMPLES], -1)
  a = lsm.least_square_mc(s2, [1, 2, 3], put, basis, _LS_DF, dtype=np.float64)
  b = lsm.least_square_mc(_LS_SAMPLES, [1, 2, 3], put, basis, _LS_DF, dtype=np.float64)
  assert a.shape == (3,)
  np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)


def test_philox_uniform_conversions():
  # random_distributions.h Uint32ToFloat / Uint64ToDouble: [0, 1), extremes exact
  from oracle import philox as ophilox
  w = np.array([[0, 0, 0xFFFFFFFF, 0xFFFFFFFF], [0x00800000, 0x007FFFFF, 0x000FFFFF, 0]], np.uint32)
  f = ophilox.uniforms_from_words(w, np.float32)
  np.testing.assert_array_equal(f[:4], np.array([0.0, 0.0, 1 - 2.0**-23, 1 - 2.0**-23], np.float32))
  assert f[4] == 0.0 and f[5] == np.float32(1 - 2.0**-23)      # only the low 23 bits are used
  d = ophilox.uniforms_from_words(w, np.float64)
  np.testing.assert_array_equal(d, np.array([0.0, 1 - 2.0**-52, (0x7FFFFF) * 2.0**-52 + 0.0,
                                             1 - 2.0**-20]))
  u = ophilox.stateless_uniform([4096, 3], [2, 2], np.float64)
  assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01


def test_heston_closures_reference_kat():
  """heston_model_test.py:175-219: drift and volatility of the piecewise-constant
  Heston process at `times[0] = 0.1`, state `[log 100, 0.045]`."""
  from oracle import models as omodels
  pw = omodels.PiecewiseConstantFunc
  drift_fn, vol_fn = omodels.heston_closures(
      pw([0.5], [1, 1.1], np.float64), pw([0.5], [1, 0.9], np.float64),
      pw([0.3], [0.1, 0.2], np.float64), pw([0.5], [0.4, 0.6], np.float64), np.float64)
  x0 = np.array([np.log(100), 0.045])
  np.testing.assert_allclose(drift_fn(0.1, x0), [-0.0225, 0.955], rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(vol_fn(0.1, x0), [[0.21213203, 0.], [0.00848528, 0.01944222]],
                             rtol=1e-6, atol=1e-6)
  # after the jumps (t = 0.6): kappa 1.1, theta 0.9, volvol 0.2, rho 0.6
  np.testing.assert_allclose(drift_fn(0.6, x0), [-0.0225, 1.1 * (0.9 - 0.045)], rtol=1e-12)
  v = np.sqrt(0.045)
  np.testing.assert_allclose(vol_fn(0.6, x0), [[v, 0.], [0.2 * 0.6 * v, 0.2 * 0.8 * v]], rtol=1e-12)


def test_black_scholes_reference_kat_for_the_c1_sanity_check():
  """vanilla_prices_test.py:32-46: the closed form the C1 price is checked against
  (tests/test_gpu_parity.py) reproduces the reference's own known values."""
  from scipy.stats import norm
  forwards = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
  strikes = np.full(5, 3.0)
  vols = np.array([0.0001, 102.0, 2.0, 0.1, 0.4])
  d1 = (np.log(forwards / strikes) + 0.5 * vols**2) / vols
  prices = forwards * norm.cdf(d1) - strikes * norm.cdf(d1 - vols)
  np.testing.assert_allclose(
      prices, [0.0, 2.0, 2.0480684764112578, 1.0002029716043364, 2.0730313058959933], atol=1e-10)
