import os, sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tf-quant-finance_b200')
import numpy as np, torch
import tff_b200 as tff
from tff_b200 import engine
dim = 64
mv = tff.models.MultivariateGeometricBrownianMotion(dim, means=np.full(dim, 0.03, np.float32),
    volatilities=np.linspace(0.1, 0.4, dim).astype(np.float32),
    corr_matrix=(0.3 + 0.7 * np.eye(dim)).astype(np.float32), dtype=np.float32)
pay = [engine.european_call(100.0, component=-1)]
def call(i):
  x0 = 100.0 * np.ones(dim, dtype=np.float32); x0[0] += np.float32(1e-4) * i
  return tff.models.euler_sampling.price(dim, mv.drift_fn(), mv.volatility_fn(), np.array([1.0], np.float32), pay,
      num_time_steps=252, num_samples=200000, initial_state=x0, random_type=tff.math.random.RandomType.SOBOL, dtype=np.float32)
for i in range(3): call(i)
pr = cProfile.Profile(); pr.enable()
for i in range(3, 6): call(i)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
