"""Per-kernel device time of one pipelined C5 pass (torch.profiler / CUPTI):
unlike an ncu launch list the kernels run back to back with a warm L2."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tf-quant-finance_b200'))
import tff_b200 as tff
from tff_b200 import engine
from tff_b200.models import closures, utils
lsm = tff.models.longstaff_schwartz
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
r, sigma = 0.1, 1.0
times = np.linspace(0.0, 1.0, 50)
drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
spec = closures.resolve_spec(drift, vol)
all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.01), dtype=np.float64)
steps, record_slot = engine.record_plan(mask, 50)
rng = engine.RngSpec(tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0)
plan = engine.Plan(spec, all_times, steps, np.array([0.0]), rng, n, np.float64)
df = np.exp(-r * times)
put = lsm.make_basket_put_payoff([1.1], dtype=np.float64)
basis = lsm.make_polynomial_basis(3)
def one():
  paths, sums = plan.paths(record_slot, 50, 0, plan.units, exp_transform=True, column_sums=True)
  return lsm.least_square_mc(paths, np.arange(50), put, basis, discount_factors=df, dtype=np.float64,
                             column_sums=sums)
for _ in range(3): one()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  p = one()
  torch.cuda.synchronize()
print('price', p)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
  a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print('%-72s n=%4d total %9.1f us avg %7.1f us' % (k, a[0], a[1], a[1] / a[0]))
if evs:
  t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
  print('span us', t1 - t0, 'busy us', sum(a[1] for a in agg.values()))
