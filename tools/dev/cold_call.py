"""Wall time of pricing calls whose inputs change every call (nothing can be re-used from the
plan cache: tables are rebuilt on the host and uploaded) against repeated identical calls."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tf-quant-finance_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import tff_b200 as tff  # noqa: E402
from tff_b200 import engine  # noqa: E402
from tff_b200.models import closures  # noqa: E402

rt = tff.math.random.RandomType


def timeit(fn, n):
  fn(0)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(n):
    fn(i + 1)
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / n * 1e3


d, v = closures.affine_closures(0.03 - 0.1**2 / 2, 0.0, 0.1)
proc = tff.models.GenericItoProcess(1, d, v, dtype=np.float64)
pay = [engine.european_call(k, log_state=True, scale=np.exp(-0.03)) for k in (600.0, 650.0, 680.0)]
c1 = lambda x0: proc.price([1.0], pay, num_samples=100_000, initial_state=np.array([x0]),
                           random_type=rt.PSEUDO_ANTITHETIC, seed=42, time_step=0.01)
print('c1 repeated %.3f ms' % timeit(lambda i: c1(np.log(700.0)), 20))
print('c1 fresh    %.3f ms' % timeit(lambda i: c1(np.log(700.0) + 1e-12 * i), 20))
heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
pay2 = [engine.european_call(100.0, log_state=True), engine.up_and_out_call(100.0, 130.0, log_state=True)]
c2 = lambda x0: heston.price([1.0], pay2, num_samples=10_000_000, initial_state=np.array([x0, 0.04]),
                             random_type=rt.SOBOL, num_time_steps=252)
print('c2 repeated %.3f ms' % timeit(lambda i: c2(np.log(100.0)), 5))
print('c2 fresh    %.3f ms' % timeit(lambda i: c2(np.log(100.0) + 1e-12 * i), 5))
if len(sys.argv) > 1:
  import cProfile
  import pstats
  pr = cProfile.Profile()
  pr.enable()
  for i in range(5):
    c1(np.log(700.0) + 1e-9 * (i + 1))
  pr.disable()
  pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
