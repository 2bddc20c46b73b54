"""Hull-White 1F Monte-Carlo swaption pricing (config C3), host side replayed on the CPU.

`hull_white.swaption_price(use_analytic_pricing=False)` hands the fused kernel (a) the
per-step table of `TQF_MODEL_HW1F` (exact OU step + discount integral, the reference's
`dt_0 = 0` weights) and (b) one `TQF_PAYOFF_HW_SWAPTION` descriptor per swaption
(`P(t_e, T_j) = exp(k_j - G_j x)`).  With `engine.Plan` replaced by a recorder (inside the
test only) the pricer runs without a GPU up to the launch; the kernel's arithmetic
(`HullWhite1FModel::step` and the swaption payoff of `csrc/tqf_paths_kernel.cuh`) is then
restated in numpy over those tables with the oracle's draws and must reproduce the oracle's
restatement of the reference pricer (`oracle/hull_white.py`, pinned by the reference's
0.71632434) -- path by path.  The kernel itself is compared with the same oracle on the device
(`tests/test_gpu_hull_white.py`, `tests/test_gpu_baseline_shapes.py`).
"""
import numpy as np
import pytest

from oracle import draws as odraws
from oracle import hull_white as ohw
from oracle import models as omodels
from tff_b200 import engine
from tff_b200.math import piecewise
from tff_b200.models import hull_white

RT = odraws.RandomType


class _RecordedPlan:
  """Stands in for `engine.Plan`: keeps what the pricer would upload."""

  def __init__(self, spec, all_times, num_steps, x0, rng, num_samples, dtype, x0_paths=None, table=None):
    self.spec, self.rng, self.num_steps, self.num_samples = spec, rng, int(num_steps), int(num_samples)
    self.all_times = np.asarray(all_times, dtype=dtype)
    self.table = spec.coef_table(self.all_times, np.dtype(dtype))[:self.num_steps]
    self.x0 = np.asarray(x0, dtype=np.float64)

  def close(self):
    pass


def _flat_rate(t):
  return 0.01 * np.ones_like(np.asarray(t))       # analytic in t: the oracle differentiates it by a complex step


def _replay(gp, num_samples=None, seed=None):
  """Payoff of every claim on every path, as the fused kernel evaluates it."""
  plan = gp.plan
  num_samples = plan.num_samples if num_samples is None else num_samples
  z = odraws.generate_mc_normal_draws(num_normal_draws=1, num_time_steps=plan.all_times.shape[0] - 1,
                                      num_sample_paths=num_samples, batch_shape=(),
                                      random_type=RT(plan.rng.random_type.value), dtype=np.float64,
                                      seed=plan.rng.seed, skip=plan.rng.skip)
  x = np.zeros(num_samples) + plan.x0[0]
  integral = np.zeros(num_samples) + plan.x0[1]
  payoffs = [None] * len(gp.payoffs)
  for i in range(plan.table.shape[0]):
    c = plan.table[i]
    x = c[2] * z[i, :, 0] + (c[0] * x + c[1])          # HullWhite1FModel::step
    integral = c[3] * x + (integral + c[4])
    for q, p in enumerate(gp.payoffs):
      d = p.desc()
      if d.expiry_step == i + 1:
        acc = sum(d.pay_coef[j] * np.exp(-d.pay_g[j] * x + d.pay_k[j]) for j in range(d.num_payments))
        swap = np.exp(-integral) * (1.0 - acc)
        payoffs[q] = np.maximum(swap if d.is_payer else -swap, 0.0) * d.scale
  return np.stack(payoffs, axis=-1)


def _replayed_sums(gp):
  """What `tqf_plan_price` returns per claim: sum, sum of squares, non-finite count, spare."""
  v = _replay(gp)
  return np.stack([v.sum(axis=0), (v * v).sum(axis=0), np.zeros(v.shape[1]), np.zeros(v.shape[1])], axis=-1)


@pytest.mark.parametrize('case', ['reference_kat', 'batch_piecewise_vol'])
def test_swaption_tables_and_descriptors_replay(monkeypatch, case):
  monkeypatch.setattr(engine, 'Plan', _RecordedPlan)
  n, seed = 2048, [4, 2]
  if case == 'reference_kat':          # swaption_test.py:85-125
    kw = dict(expiries=np.array([1.0]), fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0]]),
              fixed_leg_daycount_fractions=0.25 * np.ones((1, 4)), fixed_leg_coupon=0.011 * np.ones((1, 4)),
              notional=100., mean_reversion=0.03, time_step=0.1)
    vol = ovol = 0.02
    is_payer = True
  else:                                # two expiries, receiver + payer, piecewise-constant volatility
    kw = dict(expiries=np.array([1.0, 2.0]),
              fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0], [2.25, 2.5, 2.75, 3.0]]),
              fixed_leg_daycount_fractions=0.25 * np.ones((2, 4)), fixed_leg_coupon=0.011 * np.ones((2, 4)),
              notional=100., mean_reversion=0.03, time_step=0.1)
    vol = piecewise.PiecewiseConstantFunc([0.5, 1.5], [0.01, 0.02, 0.015], dtype=np.float64)
    ovol = omodels.PiecewiseConstantFunc([0.5, 1.5], [0.01, 0.02, 0.015], dtype=np.float64)
    is_payer = np.array([False, True])
  gp = hull_white.swaption_price(
      floating_leg_start_times=None, floating_leg_end_times=None, floating_leg_daycount_fractions=None,
      reference_rate_fn=_flat_rate, volatility=vol, is_payer_swaption=is_payer, use_analytic_pricing=False,
      num_samples=n, random_type=RT.STATELESS_ANTITHETIC, seed=seed, dtype=np.float64, _plan_only=True, **kw)
  assert isinstance(gp.plan, _RecordedPlan) and gp.plan.spec.kind == engine._lib.MODEL_HW1F
  got = _replay(gp)
  price, want = ohw.swaption_price_mc(
      reference_rate_fn=_flat_rate, volatility=ovol, is_payer_swaption=is_payer, num_samples=n,
      random_type=RT.STATELESS_ANTITHETIC, seed=seed, dtype=np.float64, return_payoffs=True, **kw)
  want = want * 100.0                                  # the kernel folds the notional into the payoff
  assert got.shape == want.shape == (n, len(kw['expiries']))
  np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-11)
  np.testing.assert_allclose(got.mean(axis=0), price, rtol=1e-12)
  assert (got > 0).any(axis=0).all() and (got == 0).any(axis=0).all()


def test_an_expiry_on_the_time_step_grid_is_priced_at_its_expiry(monkeypatch):
  """A deliberate difference from the reference.  `swaption.py:284-288` concatenates the unique
  expiries with `tf.range(time_step, longest, time_step)` without removing duplicates; when an
  earlier expiry is an exact multiple of `time_step` (1.0 with 0.25 -- not with the 0.1 of the
  reference's tests, whose accumulated 0.1 never equals 1.0) the sampler's `keep_mask` marks the
  doubled time once (`vector_hull_white.py:1027-1030`), every later sample lands one slot early
  and the last slot of the TensorArray is never written: the later swaption is valued on zeros.
  The oracle restates that bookkeeping and returns 0; the fused pricer evaluates each claim at
  the step that lands on its expiry and gives the value the closed form confirms."""
  kw = dict(expiries=np.array([1.0, 2.0]),
            fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0], [2.25, 2.5, 2.75, 3.0]]),
            fixed_leg_daycount_fractions=0.25 * np.ones((2, 4)), fixed_leg_coupon=0.011 * np.ones((2, 4)),
            notional=100., mean_reversion=0.03, reference_rate_fn=_flat_rate, volatility=0.02, dtype=np.float64)
  legs = dict(floating_leg_start_times=None, floating_leg_end_times=None, floating_leg_daycount_fractions=None)
  analytic = hull_white.swaption_price(use_analytic_pricing=True, **legs, **kw)
  np.testing.assert_allclose(analytic[0], 0.71632434, rtol=1e-6)          # swaption_test.py:85-125
  n, seed = 1 << 16, [4, 2]
  oracle = ohw.swaption_price_mc(num_samples=n, time_step=0.25, random_type=RT.STATELESS_ANTITHETIC, seed=seed, **kw)
  assert oracle[1] == 0.0 and abs(oracle[0] - analytic[0]) < 2e-2
  monkeypatch.setattr(engine, 'Plan', _RecordedPlan)
  gp = hull_white.swaption_price(use_analytic_pricing=False, num_samples=n, time_step=0.25,
                                 random_type=RT.STATELESS_ANTITHETIC, seed=seed, _plan_only=True, **legs, **kw)
  assert np.sum(np.diff(gp.plan.all_times) == 0) == 1                    # the doubled 1.0
  got = _replay(gp)
  stderr = got.std(axis=0) / np.sqrt(n)
  assert np.all(np.abs(got.mean(axis=0) - analytic) < 4 * stderr + 1e-2)   # time-discretised discounting
  np.testing.assert_allclose(got.mean(axis=0)[0], oracle[0], rtol=1e-12)


# ---- bond options and caps: the whole host flow, the kernel replaced by the replay -------------
@pytest.fixture
def replayed_kernel(monkeypatch):
  from tff_b200.models.hull_white import swaption as hw_swaption
  monkeypatch.setattr(engine, 'Plan', _RecordedPlan)
  monkeypatch.setattr(hw_swaption._GridPricing, 'sums', _replayed_sums)


def test_bond_option_price_host_flow(replayed_kernel):
  # zero_coupon_bond_option_test.py:49-73 (0.02817777) and a batch with puts and a piecewise volatility
  expiries, maturities = np.array(1.0), np.array(5.0)
  strikes = np.exp(-0.01 * maturities) / np.exp(-0.01 * expiries)
  kw = dict(strikes=strikes, expiries=expiries, maturities=maturities, discount_rate_fn=_flat_rate,
            mean_reversion=0.03, volatility=0.02, num_samples=1 << 16, time_step=0.1,
            random_type=RT.STATELESS_ANTITHETIC, seed=[1, 7])
  got = hull_white.bond_option_price(use_analytic_pricing=False, dtype=np.float64, **kw)
  want = ohw.bond_option_price_mc(**kw)
  assert got.shape == want.shape == ()
  np.testing.assert_allclose(got, want, rtol=1e-11)
  np.testing.assert_allclose(got, 0.02817777, atol=5e-4)

  vol = piecewise.PiecewiseConstantFunc([0.5, 2.0], [0.01, 0.02, 0.015], dtype=np.float64)
  ovol = omodels.PiecewiseConstantFunc([0.5, 2.0], [0.01, 0.02, 0.015], dtype=np.float64)
  kw = dict(strikes=np.array([[0.95, 0.9], [0.97, 0.8]]), expiries=np.array([[1.0, 2.0], [0.5, 2.0]]),
            maturities=np.array([[5.0, 6.0], [2.5, 10.0]]), discount_rate_fn=_flat_rate, mean_reversion=0.03,
            is_call_options=np.array([[True, False], [False, True]]), num_samples=4096, time_step=0.25,
            random_type=RT.STATELESS, seed=[3, 9])
  got = hull_white.bond_option_price(use_analytic_pricing=False, volatility=vol, dtype=np.float64, **kw)
  want = ohw.bond_option_price_mc(volatility=ovol, **kw)
  assert got.shape == (2, 2)
  np.testing.assert_allclose(got, want, rtol=1e-10)


def test_cap_floor_price_host_flow(replayed_kernel):
  # cap_floor_test.py:57-83: 0.4072088281493774 +- 1e-3; the first caplet expires at t = 0
  kw = dict(strikes=0.01 * np.ones(4), expiries=np.array([0.0, 0.25, 0.5, 0.75]),
            maturities=np.array([0.25, 0.5, 0.75, 1.0]), daycount_fractions=0.25 * np.ones(4), notional=100.0,
            reference_rate_fn=_flat_rate, mean_reversion=0.03, volatility=0.02, num_samples=50_000, time_step=0.1,
            random_type=RT.STATELESS_ANTITHETIC, seed=[42, 42])
  got = hull_white.cap_floor_price(use_analytic_pricing=False, dtype=np.float64, **kw)
  want = ohw.cap_floor_price_mc(**kw)
  assert np.shape(got) == np.shape(want) == ()
  np.testing.assert_allclose(got, want, rtol=1e-10)
  np.testing.assert_allclose(got, 0.4072088281493774, rtol=1e-3, atol=1e-3)
  floor = hull_white.cap_floor_price(use_analytic_pricing=False, dtype=np.float64, **dict(kw, is_cap=False))
  np.testing.assert_allclose(floor, ohw.cap_floor_price_mc(**dict(kw, is_cap=False)), rtol=1e-10)
