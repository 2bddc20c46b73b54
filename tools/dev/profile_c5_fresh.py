"""cProfile of the fresh-plan C5 call (new plan, paths, backward induction, price)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tf-quant-finance_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import tff_b200 as tff  # noqa: E402
from tff_b200 import engine  # noqa: E402
from tff_b200.models import closures, utils  # noqa: E402

lsm = tff.models.longstaff_schwartz
n, r, sigma = 8_000_000, 0.1, 1.0
times = np.linspace(0.0, 1.0, 50)
drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
spec = closures.resolve_spec(drift, vol)
all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.01), dtype=np.float64)
nsteps, record_slot = engine.record_plan(mask, 50)
rng = engine.RngSpec(tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0)
df = np.exp(-r * times)
put = lsm.make_basket_put_payoff([1.1], dtype=np.float64)
basis = lsm.make_polynomial_basis(3)


def call(i):
  plan = engine.Plan(spec, all_times, nsteps, np.array([1e-12 * i]), rng, n, np.float64)
  try:
    paths, csums = plan.paths(record_slot, 50, 0, plan.units, exp_transform=True, column_sums=True)
    price = lsm.least_square_mc(paths, np.arange(50), put, basis, discount_factors=df, dtype=np.float64,
                                column_sums=csums)
    return float(price[0])
  finally:
    plan.close()


for i in range(4):
  call(i)
pr = cProfile.Profile()
pr.enable()
for i in range(4, 8):
  call(i)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
