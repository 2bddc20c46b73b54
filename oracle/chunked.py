"""Oracle (test infrastructure): multi-process runs of the oracle at the
BASELINE.json sizes.

The numpy oracle needs ~10 s (C2, 2^17 paths x 252 steps) to minutes (C3, 2^20
paths x 360 steps) on one core.  Every worker below computes a slice
`path_range = (lo, hi)` of the SAME run (identical draws, see
`draws.generate_mc_normal_draws`) in its own process; `run` farms the slices out
to fresh `python -m oracle.chunked` processes (no fork: safe to call from a
process that already initialised CUDA).  Nothing here is product code.
"""
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np


def heston_chunk(job):
  """C2 slice: terminal (log-spot, variance) and the running max of the log-spot.
  job = dict(scheme='euler'|'qe', lo, hi, n, steps, params, x0, random_type, seed, skip)."""
  from oracle import draws as odraws
  from oracle import euler as oeuler
  from oracle import heston_qe as oqe
  from oracle import models as omodels
  kappa, theta, volvol, rho = job['params']
  rt = odraws.RandomType[job['random_type']]
  x0 = np.asarray(job['x0'], np.float64)
  common = dict(num_samples=job['n'], random_type=rt, seed=job.get('seed'),
                skip=job.get('skip', 0), num_time_steps=job['steps'],
                path_range=(job['lo'], job['hi']), return_extrema=True)
  if job['scheme'] == 'euler':
    d, v = omodels.heston_closures(kappa, theta, volvol, rho, np.float64)
    paths, xmax, _ = oeuler.sample(2, d, v, [job['horizon']], initial_state=x0,
                                   dtype=np.float64, **common)
  else:
    paths, xmax, _ = oqe.sample_paths(kappa, theta, volvol, rho, [job['horizon']], x0,
                                      dtype=np.float64, **common)
  return paths[:, -1, :], xmax


def swaption_chunk(job):
  """C3 slice: the discounted swaption payoffs of paths [lo, hi)."""
  from oracle import draws as odraws
  from oracle import hull_white as ohw
  kw = dict(job['kwargs'])
  rate = kw.pop('flat_rate')
  kw['random_type'] = odraws.RandomType[kw['random_type']]
  _, payoff = ohw.swaption_price_mc(
      reference_rate_fn=lambda t: rate + 0 * t, return_payoffs=True,
      path_range=(job['lo'], job['hi']), **kw)
  return payoff


def gbm_log_paths_chunk(job):
  """C5 slice: log-GBM Euler paths [rows, k, 1] of units [lo, hi) (antithetic:
  the + partners of the units, then the - partners)."""
  from oracle import draws as odraws
  from oracle import euler as oeuler
  r, sigma = job['r'], job['sigma']
  return oeuler.sample(
      1, lambda t, x: (r - sigma**2 / 2) + 0 * x,
      lambda t, x: sigma * np.ones(x.shape + (1,)), np.asarray(job['times']),
      time_step=job['time_step'], num_samples=job['n'],
      initial_state=np.array([0.0]), random_type=odraws.RandomType[job['random_type']],
      seed=job['seed'], dtype=np.float64, path_range=(job['lo'], job['hi']))


_WORKERS = {'heston': heston_chunk, 'swaption': swaption_chunk,
            'gbm_log_paths': gbm_log_paths_chunk}


def slices(total, chunk):
  return [(lo, min(lo + chunk, total)) for lo in range(0, total, chunk)]


def run(worker, jobs, procs=None):
  """[worker(job) for job in jobs] computed by `procs` fresh interpreter
  processes (job i goes to process i % procs; results come back in order)."""
  procs = procs or min(len(jobs), os.cpu_count() or 1, 16)
  if procs <= 1:
    return [_WORKERS[worker](j) for j in jobs]
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, OMP_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1', MKL_NUM_THREADS='1')
  env['PYTHONPATH'] = root + os.pathsep + env.get('PYTHONPATH', '')
  with tempfile.TemporaryDirectory() as tmp:
    running = []
    for r in range(procs):
      mine = list(range(r, len(jobs), procs))
      if not mine:
        continue
      fin, fout = os.path.join(tmp, 'in%d' % r), os.path.join(tmp, 'out%d' % r)
      with open(fin, 'wb') as f:
        pickle.dump((worker, [jobs[i] for i in mine]), f)
      running.append((mine, fout, subprocess.Popen(
          [sys.executable, '-m', 'oracle.chunked', fin, fout], env=env, cwd=root)))
    out = [None] * len(jobs)
    for mine, fout, proc in running:
      if proc.wait() != 0:
        raise RuntimeError('oracle worker failed')
      with open(fout, 'rb') as f:
        for i, res in zip(mine, pickle.load(f)):
          out[i] = res
  return out


if __name__ == '__main__':
  with open(sys.argv[1], 'rb') as _f:
    _worker, _jobs = pickle.load(_f)
  _res = [_WORKERS[_worker](_j) for _j in _jobs]
  with open(sys.argv[2], 'wb') as _f:
    pickle.dump(_res, _f)
