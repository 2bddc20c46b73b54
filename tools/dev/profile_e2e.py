"""cProfile of the public pricing entry points the bench's e2e numbers go through
(C3 swaption_price at 50M paths, C1 GenericItoProcess.price): where the host time of
one call goes.  Run on a GPU box: python tools/dev/profile_e2e.py [c3|c1]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tf-quant-finance_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import tff_b200 as tff  # noqa: E402
from tff_b200.models.hull_white import swaption as swp  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'c3'
rt = tff.math.random.RandomType
if which == 'c3':
  kw = dict(expiries=np.array(1.0), fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
            fixed_leg_daycount_fractions=0.25 * np.ones(4), fixed_leg_coupon=0.011 * np.ones(4),
            mean_reversion=0.03, volatility=0.02, notional=100.0, seed=[4, 2],
            time_step=1.0 / 360, dtype=np.float64,
            floating_leg_start_times=np.array([1.0, 1.25, 1.5, 1.75]),
            floating_leg_end_times=np.array([1.25, 1.5, 1.75, 2.0]),
            floating_leg_daycount_fractions=0.25 * np.ones(4),
            reference_rate_fn=lambda t: 0.01 + 0 * t, use_analytic_pricing=False,
            num_samples=50_000_000, random_type=rt.STATELESS)
  fn = lambda: swp.swaption_price(**kw)
else:
  from tff_b200 import engine
  from tff_b200.models import closures
  d, v = closures.affine_closures(0.03 - 0.1**2 / 2, 0.0, 0.1)
  proc = tff.models.GenericItoProcess(1, d, v, dtype=np.float64)
  pay = [engine.european_call(k, log_state=True, scale=np.exp(-0.03)) for k in (600.0, 650.0, 680.0)]
  x0 = np.array([np.log(700.0)])
  fn = lambda: proc.price([1.0], pay, num_samples=100_000, initial_state=x0,
                          random_type=rt.PSEUDO_ANTITHETIC, seed=42, time_step=0.01)
for _ in range(2):
  fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
  fn()
torch.cuda.synchronize()
print('%s: %.3f ms per call' % (which, (time.perf_counter() - t0) / 3 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
  fn()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
