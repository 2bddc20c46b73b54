"""HBM write bandwidth of the path-materialising mode (MODE_PATHS) when the
normals are cheap: float32 / float64 log-space GBM Euler paths recorded at EVERY
step (time-major [k][dim][N] buffer, 4 / 8 bytes per path and recorded date).
Prints GB/s against MEASURED_PEAKS.json's copy bandwidth."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, os.path.join(ROOT, 'tf-quant-finance_b200'))
import tff_b200 as tff
from tff_b200 import engine
from tff_b200.models import closures, utils

peak = 6529.1
try:
  peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:  # pylint: disable=broad-except
  pass
for dtype, n, k, rt in ((np.float32, 16_000_000, 64, 'SOBOL'), (np.float64, 8_000_000, 64, 'SOBOL'),
                        (np.float32, 16_000_000, 64, 'STATELESS_ANTITHETIC'),
                        (np.float64, 8_000_000, 64, 'STATELESS_ANTITHETIC')):
  times = np.linspace(1.0 / k, 1.0, k).astype(dtype)
  drift, vol = closures.affine_closures(0.03 - 0.02, 0.0, 0.2)
  spec = closures.resolve_spec(drift, vol)
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=dtype(1.0 / k), dtype=dtype)
  steps, record_slot = engine.record_plan(mask, k)
  rng = engine.RngSpec(tff.math.random.RandomType[rt], [4, 2], 0)
  plan = engine.Plan(spec, all_times, steps, np.array([0.0]), rng, n, dtype)
  out = None
  for _ in range(3):
    out = plan.paths(record_slot, k)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 5
  e0.record()
  for _ in range(reps):
    out = plan.paths(record_slot, k)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / reps
  nbytes = float(n) * k * np.dtype(dtype).itemsize
  gbs = nbytes / (ms * 1e-3) / 1e9
  print('%-8s %-22s N=%d k=%d steps=%d: %.3f ms, %.1f GB written, %.0f GB/s = %.2f of %.0f GB/s copy peak'
        % (np.dtype(dtype).name, rt, n, k, steps, ms, nbytes / 1e9, gbs, gbs / peak, peak))
  plan.close()
  del out
