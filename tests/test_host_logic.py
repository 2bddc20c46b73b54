"""CPU tests of the product's host logic and of the C-ABI surface (no GPU).

The product package must not import the oracle; here the oracle is the
checker for the host-side restatements (time grids, record plan, key
derivation, direction numbers).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import philox as ophilox
from oracle import sobol as osobol
from tff_b200 import _lib
from tff_b200 import engine
from tff_b200.math import piecewise
from tff_b200.math.random import philox
from tff_b200.math.random import sobol
from tff_b200.models import utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, 'include', 'tqf.h')).read()
  declared = set(re.findall(r'\b(tqf_[a-z0-9_]+)\s*\(', header))
  assert declared, 'no declarations found'
  handle = _lib.lib()
  for name in sorted(declared):
    assert hasattr(handle, name), name
  assert declared == set(_lib.EXPORTED_SYMBOLS)
  assert handle.tqf_version() == 100


def test_struct_layouts_match_header_sizes():
  # sizes implied by include/tqf.h on LP64
  assert C.sizeof(_lib.RngDesc) == 4 + 4 + 8 + 16 + 8 + 8 + 8 + 8 + 8
  assert C.sizeof(_lib.ModelDesc) == 8 * 4 + 5 * 8
  assert C.sizeof(_lib.PayoffDesc) == 16 + 24 + 16 + 16 + 3 * 64 * 8


def test_ctypes_structs_match_the_compiled_abi():
  sizes = (C.c_int32 * 4)()
  _lib.check(_lib.lib().tqf_abi_sizes(sizes))
  assert list(sizes) == [C.sizeof(_lib.RngDesc), C.sizeof(_lib.ModelDesc),
                         C.sizeof(_lib.PayoffDesc), C.sizeof(_lib.LsmDesc)]


def test_integration_md_struct_sketch_matches_the_abi():
  # the `_Rng` / `_Model` classes INTEGRATION.md shows a maintainer are executed as written
  text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
  ns = {'C': C}
  for cls in ('_Rng', '_Model'):
    m = re.search(r'^class %s\(C\.Structure\):\n(?:  .*\n)+' % cls, text, re.M)
    assert m, cls
    exec(m.group(0), ns)  # pylint: disable=exec-used
  sizes = (C.c_int32 * 4)()
  _lib.check(_lib.lib().tqf_abi_sizes(sizes))
  assert [C.sizeof(ns['_Rng']), C.sizeof(ns['_Model'])] == list(sizes)[:2]
  assert [f[0] for f in ns['_Rng']._fields_] == [f[0] for f in _lib.RngDesc._fields_]
  assert [f[0] for f in ns['_Model']._fields_] == [f[0] for f in _lib.ModelDesc._fields_]


def test_product_does_not_import_oracle():
  pkg = os.path.join(ROOT, 'tf-quant-finance_b200')
  for base, _, files in os.walk(pkg):
    for f in files:
      if f.endswith('.py'):
        src = open(os.path.join(base, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f


GRID_CASES = [
    dict(times=[1.0], time_step=0.01),
    dict(times=[1.0], num_time_steps=252),
    dict(times=[0.1, 0.5, 1.0, 2.0], time_step=0.1),
    dict(times=[0.0, 0.5, 1.0], time_step=0.25),
    dict(times=[1.0], time_step=1 / 252),
    dict(times=[1.0], time_step=1 / 360),
    dict(times=[0.3, 0.77, 1.9], time_step=0.07),
    dict(times=[0.3, 0.77, 1.9], num_time_steps=5),
    dict(times=[0.3, 0.77, 1.9], num_time_steps=2),
    dict(times=[0.3, 0.75], times_grid=[0.0, 0.25, 0.5, 0.75, 1.0]),
    dict(times=[0.31, 0.74], times_grid=[0.0, 0.25, 0.5, 0.75, 1.0]),
    dict(times=np.linspace(0, 1, 50), time_step=0.01),
]


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('case', GRID_CASES)
def test_prepare_grid_matches_oracle(case, dtype):
  times = np.asarray(case['times'], dtype=dtype)
  ts = case.get('time_step')
  nts = case.get('num_time_steps')
  tg = case.get('times_grid')
  if ts is None and tg is None:
    ts = dtype(times[-1] / dtype(nts))
  got = utils.prepare_grid(times=times, time_step=ts, dtype=dtype,
                           num_time_steps=nts, times_grid=tg)
  want = ogrid.prepare_grid(times=times, time_step=ts, dtype=dtype,
                            num_time_steps=nts, times_grid=tg)
  for g, w in zip(got, want):
    np.testing.assert_array_equal(g, w)
  assert got[0].dtype == dtype


def test_record_plan_replays_the_while_loop():
  # keep_mask, k -> (steps executed, slots)
  n, slots = engine.record_plan([False, False, True, False, True], 2)
  assert n == 4 and slots.tolist() == [-1, -1, 0, -1, 1]
  # times[0] == 0: the initial state is slot 0
  n, slots = engine.record_plan([True, False, True], 2)
  assert n == 2 and slots.tolist() == [0, -1, 1]
  # grid longer than the requested times: stop early
  n, slots = engine.record_plan([False, True, False, False], 1)
  assert n == 1 and slots.tolist() == [-1, 0]
  # only the initial time requested
  n, slots = engine.record_plan([True, False], 1)
  assert n == 0 and slots.tolist() == [0]


@pytest.mark.parametrize('seed', [[4, 2], [0, 0], [-1, 7], [2**31 - 1, -2**31],
                                  [123456789012, 5]])
def test_stateless_key_counter_matches_oracle(seed):
  key, ctr = philox.stateless_key_counter(seed)
  okey, octr = ophilox.stateless_key_counter(seed)
  assert list(key) == okey.tolist() and list(ctr) == octr.tolist()


@pytest.mark.parametrize('seed', [42, 0, 1, 2**31 - 1, 2**31 + 5, 87654321])
def test_stateful_key_counter_matches_oracle(seed):
  key, ctr = philox.stateful_key_counter(seed)
  okey, octr = ophilox.stateful_key_counter(seed)
  assert list(key) == okey.tolist() and list(ctr) == octr.tolist()


def test_direction_numbers_match_oracle():
  dn = sobol.direction_numbers(700)
  want = osobol.direction_numbers(700)
  # columns 0..30 are the ones `sample` can use; column 31 wraps in int32
  np.testing.assert_array_equal(dn[:, :31].astype(np.int64), want[:, :31])
  np.testing.assert_array_equal(dn.view(np.uint32)[:, 31].astype(np.int64),
                                want[:, 31] & 0xFFFFFFFF)


@pytest.mark.skipif(
    not os.path.exists('/root/reference/third_party/sobol_data/new-joe-kuo-6.21201'),
    reason='reference checkout absent')
def test_direction_numbers_from_reference_text_file():
  path = '/root/reference/third_party/sobol_data/new-joe-kuo-6.21201'
  a = sobol.direction_numbers_from_file(path, 300)
  np.testing.assert_array_equal(a, sobol.direction_numbers(300))


def test_piecewise_constant_func():
  # math/piecewise.py docstring example (lines 38-50)
  f = piecewise.PiecewiseConstantFunc([0.1, 10], [3, 4, 5], dtype=np.float64)
  np.testing.assert_array_equal(f(np.array([0., 0.1, 2., 11.])), [3, 3, 4, 5])
  x = np.array([0., 0.1, 2., 11.])
  np.testing.assert_allclose(f.integrate(x, x + 1), [3.9, 4, 4, 5])


def test_heston_coefficient_table():
  volvol = piecewise.PiecewiseConstantFunc([0.5], [1.0, 1.1], dtype=np.float64)
  spec = engine.HestonEulerSpec(0.5, 0.04, volvol, 0.1)
  t = np.array([0.0, 0.25, 0.5, 0.75])
  tab = spec.coef_table(t, np.float64)
  assert tab.shape == (3, 6)
  np.testing.assert_allclose(tab[:, 0], 0.5)
  np.testing.assert_allclose(tab[:, 4], [0.05, 0.05, 0.055])   # volvol(t_{i+1}) rho sqrt_dt


def test_no_gpu_fails_loudly():
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  import tff_b200 as tff
  with pytest.raises(_lib.TqfError):
    tff.math.random.stateless_normal([4], [1, 2], np.float64)


def test_hw_bond_price_kat_host():
  # models/hull_white/hull_white_test.py:465-485 through the product's host code
  import tff_b200 as tff
  m = tff.models.HullWhiteModel1F(0.1, 0.01, lambda t: 0.01 + 0 * t, dtype=np.float64)
  got = m.discount_bond_price([[0.011], [0.01]], [1.0, 2.0], [2.0, 3.5])
  np.testing.assert_allclose(got[:, 0], [0.98906753, 0.98495442], atol=5e-9)


def test_hw_exact_tables_match_oracle():
  from oracle import hull_white as ohw
  from oracle import models as omodels
  from tff_b200.models.hull_white import _exact
  vol = piecewise.PiecewiseConstantFunc([0.1, 0.7], [0.01, 0.02, 0.015], dtype=np.float64)
  ovol = omodels.PiecewiseConstantFunc([0.1, 0.7], [0.01, 0.02, 0.015], dtype=np.float64)
  tab = _exact.ExactTables(0.1, vol, np.float64)
  om = ohw.HullWhiteModel1F(0.1, ovol, lambda t: 0.01 + 0 * t)
  t = np.array([0.0, 0.05, 0.1, 0.3, 0.7, 0.7, 1.0, 2.5])
  np.testing.assert_array_equal(tab.conditional_mean_x(t), om.conditional_mean_x(t))
  np.testing.assert_array_equal(tab.conditional_variance_x(t), om.conditional_variance_x(t))
  np.testing.assert_array_equal(tab.y_t(t), om.compute_yt(t))
  fwd, fwd_grad = _exact.forward_rate_fns(lambda t: 0.01 + 0.002 * t, np.float64)
  np.testing.assert_allclose(fwd(np.array([0.0, 1.0])), [0.01, 0.014], rtol=1e-14)
  np.testing.assert_allclose(fwd_grad(np.array([0.0, 1.0])), [0.004, 0.004], rtol=1e-9)


# ----- host tables of the generators / model specs added for SURVEY 8f (no GPU needed)
def test_halton_host_tables_equal_the_oracle():
  from oracle import halton as ohalton
  from tff_b200.math.random import halton
  for dtype in (np.float32, np.float64):
    for dim in (1, 2, 40, 1000):
      radixes, sizes, weights, max_size = halton._tables(dim, np.dtype(dtype))
      np.testing.assert_array_equal(radixes, ohalton.primes(dim))
      np.testing.assert_array_equal(sizes, ohalton.max_sizes_by_axes(dim, dtype).reshape(-1).astype(np.int32))
      assert max_size == int(sizes.max()) and weights.shape == (dim, max_size)
      # weights = round(radix ** j) in dtype, exactly representable, 1 beyond an axis' own digits
      for d in (0, dim - 1):
        p = int(radixes[d])
        np.testing.assert_array_equal(weights[d, :sizes[d]], [float(p**j) for j in range(sizes[d])])
        assert np.all(weights[d, sizes[d]:] == 1.0)
  assert halton._first_primes(1000)[-1] == 7919
  with pytest.raises(ValueError):
    halton.sample(3, randomized=False)
  with pytest.raises(NotImplementedError):
    halton._range_of(None, np.array([0, 2, 5]))           # non-contiguous indices


def test_milstein_and_tangent_specs_build_the_documented_tables():
  from tff_b200 import engine
  all_times = np.array([0.0, 0.25, 0.5, 1.0])
  gbm = engine.GbmSpec1F(0.05, lambda t: 0.2 + 0.1 * np.asarray(t))
  tab = engine.MilsteinSpec1F(gbm).coef_table(all_times, np.float64)
  assert tab.shape == (3, 6)
  np.testing.assert_allclose(tab[:, 0], [0.25, 0.25, 0.5])
  np.testing.assert_allclose(tab[:, 1], np.sqrt([0.25, 0.25, 0.5]))
  np.testing.assert_array_equal(tab[:, 2], 0.0)                       # a0
  np.testing.assert_allclose(tab[:, 3], 0.05)                         # a1 = mu
  np.testing.assert_array_equal(tab[:, 4], 0.0)                       # b0
  np.testing.assert_allclose(tab[:, 5], [0.225, 0.25, 0.3])           # b1 = sigma(times[i + 1])
  aff = engine.AffineSpec1F(0.1, -0.2, 0.3, 0.4)
  np.testing.assert_array_equal(engine.MilsteinSpec1F(aff).coef_table(all_times, np.float64),
                                aff.coef_table(all_times, np.float64))
  with pytest.raises(NotImplementedError):
    engine.MilsteinSpec1F(engine.HestonEulerSpec(2.0, 0.04, 0.5, -0.7))
  tan = engine.TangentAffineSpec1F(0.1, -0.2, 0.3, 0.4, da0=1.0, db=2.0)
  tab = tan.coef_table(all_times, np.float64)
  assert tab.shape == (3, 10) and tan.dim == 3 and tan.user_dim == 1
  np.testing.assert_allclose(tab[0], [0.25, 0.5, 0.1, -0.2, 0.3, 0.4, 1.0, 0.0, 2.0, 0.0])
  np.testing.assert_array_equal(tan.extend_initial_state(np.array([1.5])), [1.5, 1.0, 0.0])


def test_uniform_and_milstein_argument_errors_need_no_gpu():
  import tff_b200 as tff
  rt = tff.math.random.RandomType
  with pytest.raises(ValueError):
    tff.math.random.uniform(2, [4], random_type=rt.STATELESS)                 # no seed
  with pytest.raises(NotImplementedError):
    tff.math.random.uniform(2, [4], random_type=rt.PSEUDO_ANTITHETIC, seed=1)
  from tff_b200.models import closures
  drift, vol = closures.gbm_closures(0.1, 0.2)
  with pytest.raises(NotImplementedError):
    tff.models.milstein_sampling.sample(dim=2, drift_fn=drift, volatility_fn=vol, times=[1.0],
                                        time_step=0.1)
  with pytest.raises(ValueError):
    tff.models.milstein_sampling.sample(dim=1, drift_fn=drift, volatility_fn=vol, times=[1.0])


# ProbedAffineSpec (engine.py): plain Python callables are accepted only when they
# are affine on the WHOLE state space; the probing runs on the host.
@pytest.mark.parametrize('name', ['relu', 'abs', 'clamp', 'cutoff', 'reciprocal', 'log'])
def test_probed_affine_spec_rejects_piecewise_and_singular_callables(name):
  import torch
  from tff_b200 import engine
  drifts = {
      'relu': lambda t, x: torch.relu(x),
      'abs': lambda t, x: torch.abs(x),
      'clamp': lambda t, x: torch.clamp(x, -50.0, 50.0),
      'cutoff': lambda t, x: torch.where(x > 130.0, torch.zeros_like(x), 0.1 * x),
      'reciprocal': lambda t, x: 1.0 / x,
      'log': lambda t, x: torch.log(x),
  }
  vol = lambda t, x: 0.2 * torch.ones_like(x).unsqueeze(-1)
  spec = engine.ProbedAffineSpec(1, drifts[name], vol)
  spec.initial_state_hint = np.array([100.0])
  with pytest.raises(NotImplementedError):
    spec.coef_table(np.linspace(0.0, 1.0, 5), np.float64)


def test_probed_affine_spec_accepts_affine_callables():
  import torch
  from tff_b200 import engine
  spec = engine.ProbedAffineSpec(
      2, lambda t, x: torch.stack([0.1 * t + 0.5 * x[..., 1], -0.3 * x[..., 0] + 1.0], -1),
      lambda t, x: torch.eye(2, dtype=x.dtype).expand(x.shape[0], 2, 2) * (1.0 + t))
  spec.initial_state_hint = np.array([1.0, -2.0])
  tab = spec.coef_table(np.linspace(0.0, 1.0, 5), np.float64)
  assert tab.shape == (4, spec.num_coef)
  np.testing.assert_allclose(tab[:, 2:4], np.stack([0.1 * np.linspace(0.25, 1, 4), np.ones(4)], -1))
  np.testing.assert_allclose(tab[0, 4:8], [0.0, 0.5, -0.3, 0.0], atol=1e-12)


# Closed-form Hull-White valuations (host): the reference's analytic known answers
def _flat_rate(t):
  import torch
  return 0.01 + 0 * t if isinstance(t, torch.Tensor) else 0.01 * np.ones_like(np.asarray(t, dtype=np.float64))


def test_analytic_swaption_reference_kats():
  # hull_white/swaption_test.py:85-125 (0.71632434), :160-205 (time-dependent volatility,
  # 0.5593057004094042), :127-158 (receiver, 0.813482544626056).  The reference's values
  # come out of a Brent search stopped at its default root tolerance of 2e-7
  # (math/root_search/brent.py), and the price moves by ~30 per unit of break-even rate:
  # they are good to ~1e-6, which is the tolerance here (the root below is solved to 1e-14).
  import tff_b200 as tff
  from tff_b200.math import piecewise
  kw = dict(expiries=np.array(1.0), floating_leg_start_times=np.array([1.0, 1.25, 1.5, 1.75]),
            floating_leg_end_times=np.array([1.25, 1.5, 1.75, 2.0]),
            fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
            floating_leg_daycount_fractions=0.25 * np.ones(4),
            fixed_leg_daycount_fractions=0.25 * np.ones(4), fixed_leg_coupon=0.011 * np.ones(4),
            reference_rate_fn=_flat_rate, notional=100., mean_reversion=0.03, dtype=np.float64)
  price = tff.models.hull_white.swaption_price(volatility=0.02, **kw)
  assert price.shape == () and price.dtype == np.float64
  np.testing.assert_allclose(price, 0.71632434, rtol=0, atol=2e-6)
  price = tff.models.hull_white.swaption_price(volatility=0.02, is_payer_swaption=False, **kw)
  np.testing.assert_allclose(price, 0.813482544626056, rtol=0, atol=2e-6)
  vol = piecewise.PiecewiseConstantFunc([0.5], [0.01, 0.02], dtype=np.float64)
  price = tff.models.hull_white.swaption_price(volatility=vol, **kw)
  np.testing.assert_allclose(price, 0.5593057004094042, rtol=0, atol=2e-6)


def test_analytic_bond_option_reference_kat():
  # hull_white/zero_coupon_bond_option_test.py:49-73
  import tff_b200 as tff
  expiries, maturities = np.array(1.0), np.array(5.0)
  strikes = np.exp(-0.01 * maturities) / np.exp(-0.01 * expiries)
  price = tff.models.hull_white.bond_option_price(
      strikes=strikes, expiries=expiries, maturities=maturities, mean_reversion=0.03,
      volatility=0.02, discount_rate_fn=_flat_rate, dtype=np.float64)
  assert price.shape == ()
  np.testing.assert_allclose(price, 0.02817777, rtol=1e-8, atol=1e-8)


def test_heston_model_closures_reference_kat():
  """heston_model_test.py:175-219 on the mirror's own `drift_fn()` / `volatility_fn()`
  (they evaluate on the host; the device kernel gets the same numbers as a coefficient table)."""
  import tff_b200 as tff
  pw = piecewise.PiecewiseConstantFunc
  process = tff.models.HestonModel(
      mean_reversion=pw([0.5], [1, 1.1], dtype=np.float64), theta=pw([0.5], [1, 0.9], dtype=np.float64),
      volvol=pw([0.3], [0.1, 0.2], dtype=np.float64), rho=pw([0.5], [0.4, 0.6], dtype=np.float64),
      dtype=np.float64)
  x0 = np.array([np.log(100), 0.045])
  np.testing.assert_allclose(np.asarray(process.drift_fn()(0.1, x0)), [-0.0225, 0.955],
                             rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(np.asarray(process.volatility_fn()(0.1, x0)),
                             [[0.21213203, 0.], [0.00848528, 0.01944222]], rtol=1e-6, atol=1e-6)
  # the coefficient table the kernel reads holds the same parameters, taken at t_{i+1}
  # (`euler_sampling.py:519`): columns sqrt(dt), -dt/2, dt kappa, theta, ...
  tab = process.drift_fn().tqf_spec.coef_table(np.array([0.0, 0.1, 0.6, 0.7]), np.float64)
  np.testing.assert_allclose(tab[:, 2], [0.1 * 1.0, 0.5 * 1.1, 0.1 * 1.1])
  np.testing.assert_allclose(tab[:, 3], [1.0, 0.9, 0.9])
