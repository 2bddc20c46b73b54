"""cProfile of fresh-parameter HestonModel.price calls (C2 shape at 200k paths): the host work
of one pricing call when nothing comes from the plan cache."""
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
from tff_b200 import engine  # noqa: E402

heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
pay = [engine.european_call(100.0, log_state=True), engine.up_and_out_call(100.0, 130.0, log_state=True)]
call = lambda i: heston.price([1.0], pay, num_samples=200_000, num_time_steps=252,
                              initial_state=np.array([np.log(100.0) + 1e-12 * i, 0.04]),
                              random_type=tff.math.random.RandomType.SOBOL)
for i in range(12):
  call(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(12, 32):
  call(i)
print('fresh call: %.3f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for i in range(32, 52):
  call(i)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
