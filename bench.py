#!/usr/bin/env python
"""Benchmark of the fused Euler Monte-Carlo path engine (BASELINE.json metric:
Euler path-steps/sec, fp64).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

One "step" = one pass of the hot path over the whole workload (all paths x all
Euler steps, normals generated in-kernel, payoffs reduced in-kernel).  Default
workload = BASELINE.json configs[1] (C2): Heston Euler, 10M paths x 252 steps,
float64, Sobol, European + up-and-out barrier call.  Paths shard across ranks
by disjoint Sobol index ranges / Philox counter ranges; the only collective is
the all-reduce of the per-GPU payoff sums ("scaling": "weak" is not used: the
total workload is the named config, so scaling is "strong").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'tf-quant-finance_b200')
for _p in (ROOT, PKG):
  if _p not in sys.path:
    sys.path.insert(0, _p)

METRIC = 'euler_path_steps_per_sec'
UNIT = 'path-steps/s'

# Algorithmic FP64-pipe instructions per path-step (DESIGN.md section 4.1,
# SURVEY.md 8(d); frozen in roofline.json).  C2 = 2 Sobol normals x 32 (t, 1-t^2,
# table log 8, degree-21 polynomial, t*P) + Heston Euler update 14 (sqrt 6, state 8)
# + barrier compare 1; it was 96 before the table logarithm replaced the
# 18-instruction log.  C3 / C1 were 47 / 26 (SURVEY's budget of 44 per Philox +
# Box-Muller normal); the hand-written log / sqrt / sincos need 25 per normal (a
# Box-Muller evaluation of ~50 yields TWO normals), so A = 25 + 4 (C3) and
# 25 / 2 + 3.5 (C1, antithetic: one normal serves two paths): SURVEY allows A to be
# tightened only downward, and ncu shows the C3 kernel executing 29.7 per path-step.
ALGO_FP64_INSTR = {'c1': 16, 'c2': 79, 'c3': 29, 'c4': 3616, 'c5': 22}

WORKLOADS = {
    'c1': dict(name='C1 GBM call (log-space affine), 100k paths x 100 steps, fp64, PSEUDO_ANTITHETIC seed 42',
               paths=100_000, steps=100, dtype='f64'),
    'c2': dict(name='C2 Heston Euler, 10M paths x 252 steps, fp64, Sobol, European + up-and-out call',
               paths=10_000_000, steps=252, dtype='f64'),
    'c3': dict(name='C3 Hull-White 1F payer swaption (exact OU step + discount integral), 50M paths x 360 steps, fp64, Philox stateless seed [4,2]',
               paths=50_000_000, steps=360, dtype='f64'),
    'c4': dict(name='C4 correlated 64-asset GBM basket call, 20M paths x 252 steps, fp32, Sobol + Cholesky',
               paths=20_000_000, steps=252, dtype='f32'),
    'c5': dict(name='C5 American put, Longstaff-Schwartz on log-GBM Euler paths, 8M paths x 50 exercise dates '
                    '(148 Euler steps, time_step 0.01), fp64, STATELESS_ANTITHETIC seed [4,2], cubic basis',
               paths=8_000_000, steps=148, dtype='f64'),
}


# ------------------------------------------------------------ workloads ----
def make_workload(name, num_paths=None):
  """Returns (spec, all_times, x0, rng kwargs, payoffs, paths, steps)."""
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  from tff_b200.models import utils
  w = WORKLOADS[name]
  n = int(num_paths or w['paths'])
  rt = tff.math.random.RandomType
  if name == 'c2':
    model = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5,
                                   rho=-0.7, dtype=np.float64)
    spec = closures.resolve_spec(model.drift_fn(), model.volatility_fn())
    times = np.array([1.0])
    all_times, mask, _ = utils.prepare_grid(
        times=times, time_step=np.float64(1.0 / 252), num_time_steps=252,
        dtype=np.float64)
    x0 = np.array([np.log(100.0), 0.04])
    rng = dict(random_type=rt.SOBOL, seed=None, skip=0)
    payoffs = [engine.european_call(100.0, log_state=True),
               engine.up_and_out_call(100.0, 130.0, log_state=True)]
  elif name == 'c1':
    r, sigma = 0.03, 0.1
    d, v = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
    spec = closures.resolve_spec(d, v)
    times = np.array([1.0])
    all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.01),
                                            dtype=np.float64)
    x0 = np.array([np.log(700.0)])
    rng = dict(random_type=rt.PSEUDO_ANTITHETIC, seed=42, skip=0)
    payoffs = [engine.european_call(k, log_state=True, scale=np.exp(-r))
               for k in (600.0, 650.0, 680.0)]
  elif name == 'c3':
    # swaption_test.py:30-44, 81-125 scaled up: 1y x 1y payer swaption, quarterly
    # payments, a = 0.03, sigma = 0.02, flat 1% curve, time_step 1/360.
    from tff_b200.models.hull_white import one_factor
    from tff_b200.models.hull_white import swaption as swp
    model = one_factor.HullWhiteModel1F(0.03, 0.02, lambda t: 0.01 + 0 * t,
                                        dtype=np.float64)
    ts = np.float64(1.0 / 360)
    sim_times = np.sort(np.concatenate(
        [[1.0], utils._tf_range(ts, 1.0, ts, np.float64)]), kind='stable')
    all_times, mask, idx = model._prepare_grid(sim_times, None)
    dts = np.concatenate([[0.0], sim_times[1:] - sim_times[:-1]])
    w = np.zeros(all_times.shape[0] - 1)
    for j, i in enumerate(idx):
      if i >= 1:
        w[i - 1] += dts[j]
    spec = one_factor.HullWhite1FSpec(model._tables, model._fwd, w)
    x0 = np.zeros(2)
    rng = dict(random_type=rt.STATELESS, seed=[4, 2], skip=0)
    pay = np.array([1.25, 1.5, 1.75, 2.0])
    e_idx = idx[np.searchsorted(sim_times, 1.0, side='left')]
    payoffs = [swp._RawPayoff(swp._swaption_desc(
        model, e_idx, 1.0, pay, 0.011 * np.ones(4), 0.25 * np.ones(4),
        True, 100.0))]
    steps = int(e_idx)
    return spec, all_times, x0, rng, payoffs, n, steps
  elif name == 'c4':
    dim = 64
    spec = engine.MvGbmSpec(np.full(dim, 0.03, np.float32),
                            np.linspace(0.1, 0.4, dim).astype(np.float32),
                            (0.3 + 0.7 * np.eye(dim)).astype(np.float32), dim)
    times = np.array([1.0], dtype=np.float32)
    all_times, mask, _ = utils.prepare_grid(
        times=times, time_step=np.float32(1.0) / np.float32(252), num_time_steps=252,
        dtype=np.float32)
    x0 = 100.0 * np.ones(dim, dtype=np.float32)
    rng = dict(random_type=rt.SOBOL, seed=None, skip=0)
    payoffs = [engine.european_call(100.0, component=-1)]
  else:
    raise ValueError(name)
  steps, _ = engine.record_plan(mask, 1)
  return spec, all_times, x0, rng, payoffs, n, steps


# ---------------------------------------------------------------- clocks ----
class ClockSampler(threading.Thread):
  """Samples nvidia-smi clocks / throttle reasons during the timed region."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self._halt = index, [], set(), False
    self.max_mhz = None

  def run(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    while not self._halt:
      try:
        out = subprocess.run(
            ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
             '--format=csv,noheader,nounits'], capture_output=True, text=True,
            timeout=5).stdout.strip().split(',')
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for nm, val in zip(names, out[2:]):
          if val.strip().lower().startswith('active'):
            self.reasons.add(nm)
      except Exception:  # pylint: disable=broad-except
        pass
      time.sleep(0.1)

  def stop(self):
    self._halt = True
    self.join(timeout=3)
    med = float(np.median(self.samples)) if self.samples else None
    return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# --------------------------------------------------------- CPU baselines ----
def _oracle_chunk(args):
  name, skip, count = args
  from oracle import draws as odraws
  from oracle import euler as oeuler
  from oracle import models as omodels
  if name == 'c2':
    d, v = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, np.float64)
    paths = oeuler.sample(2, d, v, [1.0], num_time_steps=252, num_samples=count,
                          initial_state=np.array([np.log(100.0), 0.04]),
                          random_type=odraws.RandomType.SOBOL, skip=skip,
                          dtype=np.float64)
    st = np.exp(paths[:, -1, 0])
    return float(np.maximum(st - 100.0, 0).sum()), 252
  if name == 'c1':
    r, sigma = 0.03, 0.1
    paths = oeuler.sample(
        1, lambda t, x: (r - sigma**2 / 2) + 0 * x,
        lambda t, x: sigma * np.ones(x.shape + (1,)), [1.0], time_step=0.01,
        num_samples=count, initial_state=np.array([np.log(700.0)]),
        random_type=odraws.RandomType.PSEUDO_ANTITHETIC, seed=42 + skip,
        dtype=np.float64)
    return float(np.maximum(np.exp(paths[:, 0, 0]) - 650.0, 0).sum()), 100
  if name == 'c3':
    from oracle import hull_white as ohw
    price, payoff = ohw.swaption_price_mc(
        expiries=np.array(1.0), fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
        fixed_leg_daycount_fractions=0.25 * np.ones(4),
        fixed_leg_coupon=0.011 * np.ones(4), reference_rate_fn=lambda t: 0.01 + 0 * t,
        mean_reversion=0.03, volatility=0.02, notional=100., num_samples=count,
        random_type=odraws.RandomType.STATELESS, seed=[4, 2 + skip],
        time_step=1.0 / 360, dtype=np.float64, return_payoffs=True)
    return float(payoff.sum()), 360
  if name == 'c4':
    dim = 64
    d, v = omodels.mvgbm_closures(np.full(dim, 0.03, np.float32),
                                  np.linspace(0.1, 0.4, dim).astype(np.float32),
                                  (0.3 + 0.7 * np.eye(dim)).astype(np.float32), np.float32)
    paths = oeuler.sample(dim, d, v, np.array([1.0], np.float32), num_time_steps=252,
                          num_samples=count, initial_state=100.0 * np.ones(dim, np.float32),
                          random_type=odraws.RandomType.SOBOL, skip=skip, dtype=np.float32)
    return float(np.maximum(paths[:, 0, :].mean(axis=1) - 100.0, 0).sum()), 252
  if name == 'c5':
    from oracle import lsm as olsm
    r, sigma = 0.1, 1.0
    times = np.linspace(0.0, 1.0, 50)
    paths = np.exp(oeuler.sample(
        1, lambda t, x: (r - sigma**2 / 2) + 0 * x,
        lambda t, x: sigma * np.ones(x.shape + (1,)), times, time_step=0.01,
        num_samples=count, initial_state=np.array([0.0]),
        random_type=odraws.RandomType.STATELESS_ANTITHETIC, seed=[4, 2 + skip],
        dtype=np.float64))
    price = olsm.least_square_mc(paths, np.arange(50), olsm.make_basket_put_payoff([1.1]),
                                 olsm.make_polynomial_basis(3), np.exp(-r * times),
                                 dtype=np.float64)
    return float(price[0]) * count, 148
  raise ValueError(name)


def cpu_run(name, sample_paths, procs):
  """Times the oracle (numpy port of the reference path: precomputed draws
  tensor + one vectorised update per step) on `sample_paths` paths split over
  `procs` worker processes.  Returns (path_steps_per_s, seconds, steps)."""
  import multiprocessing as mp
  chunk = max(sample_paths // procs, 2)
  chunk -= chunk % 2
  jobs = [(name, i * chunk, chunk) for i in range(procs)]
  t0 = time.perf_counter()
  if procs == 1:
    res = [_oracle_chunk(jobs[0])]
  else:
    with mp.get_context('fork').Pool(procs) as pool:
      res = pool.map(_oracle_chunk, jobs)
  dt = time.perf_counter() - t0
  steps = res[0][1]
  return chunk * procs * steps / dt, dt, steps, chunk * procs


def run_reference(args):
  """`--impl reference`: the reference's CPU path.  TensorFlow cannot be
  installed in this image (no wheel, no network), so the oracle port -- the
  numpy restatement pinned by the reference's known-answer tests -- is timed
  on all host cores."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  sample = {'c1': 100_000, 'c2': 8192 * cores, 'c3': 16384 * cores, 'c4': 256 * cores,
            'c5': 16384 * cores}[args.workload]
  for _ in range(args.warmup):
    cpu_run(args.workload, max(sample // 8, 2 * cores), cores)
  vals, secs = [], []
  for _ in range(args.steps):
    v, dt, steps, n = cpu_run(args.workload, sample, cores)
    vals.append(v)
    secs.append(dt)
  value = float(np.sum([sample * steps for _ in secs]) / np.sum(secs))
  w = WORKLOADS[args.workload]
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * float(np.mean(secs)), 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': w['dtype'],
      'data': 'synthetic',
      'config': {'workload': w['name'], 'sample_paths': n,
                 'note': 'bounded sample of the workload; oracle port of the TF CPU path'},
      'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                       'sample': '%d paths x %d steps per step, %d processes' % (n, steps, cores)},
      'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  emit(line)


# -------------------------------------------------------------- GPU arm ----
def run_gpu(args):
  import torch
  import torch.distributed as dist
  from tff_b200 import engine

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

  spec, all_times, x0, rngkw, payoffs, n, steps = make_workload(args.workload, args.paths)
  rng = engine.RngSpec(**rngkw)
  wdtype = np.float32 if WORKLOADS[args.workload]['dtype'] == 'f32' else np.float64
  plan = engine.Plan(spec, all_times, steps, x0, rng, n, wdtype)
  units = plan.units
  per = (units + world - 1) // world
  lo, hi = min(rank * per, units), min((rank + 1) * per, units)
  stream = torch.cuda.current_stream()

  # several GPUs: the payoff sums of the ranks are added inside the reduction
  # kernel over NVLink peer memory (no NCCL call in the step);
  # TQF_PRICE_PEER_EXCHANGE=0 selects the NCCL all-reduce instead
  px = None
  if world > 1 and os.environ.get('TQF_PRICE_PEER_EXCHANGE', '1') != '0':
    from tff_b200 import distributed
    try:
      px = distributed.PeerExchange()       # fails on ALL ranks together or on none
      plan.set_peer_exchange(px)
    except RuntimeError as e:
      sys.stderr.write('peer exchange unavailable (%s): NCCL all-reduce of the sums\n' % e)
      px = None

  def one_step():
    sums = plan.price_sums(payoffs, lo, hi - lo)
    if world > 1 and px is None:
      dist.all_reduce(sums)
    return sums

  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
  for _ in range(max(args.warmup, 3)):
    sums = one_step()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  sampler = ClockSampler(local) if rank == 0 else None
  if sampler:
    sampler.start()
  evs = []
  torch.cuda.synchronize()
  t_wall0 = time.perf_counter()
  for _ in range(args.steps):
    flush.zero_()                           # evict L2 between timed iterations
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sums = one_step()
    e1.record(stream)
    evs.append((e0, e1))
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  t_wall = time.perf_counter() - t_wall0
  clocks = sampler.stop() if sampler else None
  ms = sum(a.elapsed_time(b) for a, b in evs)
  t = torch.tensor([ms], dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_total = float(t.item())
  ms_per_step = ms_total / args.steps
  value = n * steps / (ms_per_step * 1e-3)
  prices = (sums[:, 0] / n).cpu().numpy().tolist()

  # End to end through the public API with HOST buffers: model parameters and
  # times are numpy arrays (H2D of the coefficient / direction-number tables
  # happens inside), the result comes back as a numpy array (D2H of the sums).
  table_bytes = steps * spec.num_coef * 8 + 8 * spec.dim
  if rng.type == 2:
    table_bytes += plan.num_steps_total * spec.num_factors * 32 * 4
  d2h_bytes = len(payoffs) * 4 * 8

  def e2e_step():
    p = engine.Plan(spec, all_times, steps, x0, engine.RngSpec(**rngkw), n, wdtype)
    if px is not None:
      p.set_peer_exchange(px)
    s = p.price_sums(payoffs, lo, hi - lo)
    if world > 1 and px is None:
      dist.all_reduce(s)
    out = s.cpu().numpy()[:, 0] / n
    p.close()
    return out

  e2e_step()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    e2e_step()
  torch.cuda.synchronize()
  e2e_s = time.perf_counter() - t0
  te = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
  e2e_value = n * steps * args.steps / float(te.item())

  if rank == 0:
    dfma, ffma = engine.measure_fma_peaks()
    algo = ALGO_FP64_INSTR[args.workload]
    # dominant kernel = path_kernel; its share of the step is ~100% (the
    # reduce kernel is a few microseconds) -- see profiles/.
    per_gpu_rate = (hi - lo) * (2 if rng.antithetic else 1) * steps / (ms_per_step * 1e-3)
    achieved = per_gpu_rate * algo / 1e9
    fp32 = WORKLOADS[args.workload]['dtype'] == 'f32'
    peak = (ffma if fp32 else dfma) / 1e9
    roofline = {'bound': 'fp32' if fp32 else 'fp64', 'achieved': achieved, 'peak': peak,
                'unit': 'G %s-pipe instr/s' % ('FP32' if fp32 else 'FP64'),
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full at a
                # reduced path count; the kernel only reads its tables, so the figure does not
                # grow with the path count): profiles/r1i_c2_table_log.txt, r1e_c3_hw_swaption.txt,
                # r1r_c4_mvgbm_mma.txt
                'traffic': {'c2': 196608.0, 'c3': 105984.0, 'c4': 2203392.0}.get(args.workload),
                'traffic_unit': 'bytes per launch (ncu, reduced path count)',
                'frac': achieved / peak,
                'note': 'achieved = path-steps/s/GPU x %d algorithmic %s instr per path-step; '
                        'peak = %s issue rate measured live by tqf_measure_fp64_peak '
                        '(MEASURED_PEAKS.json has no FP64/FP32 entry); kernel has no HBM traffic'
                        % (algo, 'FP32' if fp32 else 'FP64', 'FFMA' if fp32 else 'DFMA')
                        + ('; C3 is bound by the dispatch port, not by the FP64 pipe: 30 FP64 + 59 other '
                           'instructions per path-step, 47 of them the Philox rounds (roofline.json)'
                           if args.workload == 'c3' else '')}
    cores = 1
    csample = {'c1': 100_000, 'c2': 32768, 'c3': 65536, 'c4': 1024, 'c5': 65536}[args.workload]
    cv, cdt, csteps, cn = cpu_run(args.workload, csample, cores)
    w = WORKLOADS[args.workload]
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': w['dtype'], 'data': 'synthetic',
        'config': {'workload': w['name'] if args.paths is None else w['name'] + ' [paths=%d]' % n,
                   'paths': n, 'euler_steps': steps, 'payoffs': len(payoffs),
                   'sharding': 'disjoint path ranges per rank; ' + ('payoff sums added over NVLink peer memory inside the reduction kernel' if px is not None else ('NCCL all-reduce of the payoff sums' if world > 1 else 'single GPU, no exchange')),
                   'l2': 'flushed (256 MiB memset) between timed iterations; the kernel reads <100 KB of tables'},
        'prices': prices,
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': table_bytes,
                'd2h_bytes_per_step': d2h_bytes},
        'gpu_launches': 2 * args.steps,
        'wall_s_timed_region': t_wall,
        'roofline': roofline,
        'fp32_ffma_peak_ginstr': ffma / 1e9,
        'cpu_baseline': {'value': cv, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d paths x %d steps, single process numpy oracle (%.1f s)' % (cn, csteps, cdt)},
    }
    emit(line)
  plan.close()
  if px is not None:
    px.close()
  if world > 1:
    dist.destroy_process_group()


def run_gpu_c5(args):
  """C5: materialise 8M x 50 log-GBM Euler paths (time-major) and run the
  Longstaff-Schwartz passes on them.  One step = generation + regression."""
  import torch
  import torch.distributed as dist
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  lsm = tff.models.longstaff_schwartz
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  w = WORKLOADS['c5']
  n = int(args.paths or w['paths'])
  r, sigma = 0.1, 1.0
  times = np.linspace(0.0, 1.0, 50)
  drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
  spec = closures.resolve_spec(drift, vol)
  from tff_b200.models import utils
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.01),
                                          dtype=np.float64)
  steps, record_slot = engine.record_plan(mask, 50)
  rng = engine.RngSpec(tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0)
  plan = engine.Plan(spec, all_times, steps, np.array([0.0]), rng, n, np.float64)
  units = plan.units
  per = (units + world - 1) // world
  lo, hi = min(rank * per, units), min((rank + 1) * per, units)
  df = np.exp(-r * times)
  put = lsm.make_basket_put_payoff([1.1], dtype=np.float64)
  basis = lsm.make_polynomial_basis(3)
  reduce_fn = (lambda t: dist.all_reduce(t)) if world > 1 else None
  # several GPUs: the per-date normal equations are summed over the ranks inside
  # the streaming kernel through peer memory (NVLink); NCCL only carries the
  # column sums and the final value sum
  px = None
  if world > 1:
    from tff_b200 import distributed
    try:
      px = distributed.PeerExchange()       # fails on ALL ranks together or on none
    except RuntimeError as e:
      sys.stderr.write('peer exchange unavailable (%s): one NCCL all-reduce per date\n' % e)
      px = None
  stream = torch.cuda.current_stream()
  times_ms = {'gen': 0.0, 'lsm': 0.0}

  def one_step(timed):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    # [rows, 50, 1] time-major view of exp(log-price), exponentiated on store
    # (the kernel that writes the paths also sums every column: the LSM basis means)
    paths, csums = plan.paths(record_slot, 50, lo, hi - lo, exp_transform=True, column_sums=True)
    e1.record(stream)
    # antithetic shard rows: [+ partners of units lo..hi) | - partners]; the global
    # index only matters for num_calibration_samples (unused here)
    price = lsm.least_square_mc(paths, np.arange(50), put, basis, discount_factors=df,
                                dtype=np.float64, global_path_offset=2 * lo,
                                all_reduce=reduce_fn, column_sums=csums, peer_exchange=px)
    e2.record(stream)
    if timed:
      torch.cuda.synchronize()
      times_ms['gen'] += e0.elapsed_time(e1)
      times_ms['lsm'] += e1.elapsed_time(e2)
    return price

  for _ in range(max(args.warmup, 3)):
    price = one_step(False)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  sampler = ClockSampler(local) if rank == 0 else None
  if sampler:
    sampler.start()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    price = one_step(True)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  wall = time.perf_counter() - t0
  clocks = sampler.stop() if sampler else None
  tt = torch.tensor([times_ms['gen'] + times_ms['lsm'], times_ms['gen'], times_ms['lsm'], wall * 1e3],
                    dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
  tot, gen, lsm_ms, wall_ms = (float(v) / args.steps for v in tt.tolist())
  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # pylint: disable=broad-except
      pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    rows = 2 * (hi - lo)
    lsm_bytes = 49.0 * rows * 32.0
    achieved = lsm_bytes / (lsm_ms * 1e-3) / 1e9
    cv, cdt, csteps, cn = cpu_run('c5', 65536, 1)
    line = {
        'metric': METRIC, 'value': n * steps / (wall_ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': wall_ms,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': w['name'] if args.paths is None else w['name'] + ' [paths=%d]' % n,
                   'paths': n, 'euler_steps': steps, 'exercise_dates': 50,
                   'l2': 'inputs (3.2 GB of paths) exceed L2',
                   'timing': 'wall clock around generation + LSM (49 dates: one streaming pass each, the '
                             'regression solved by its last CTA; device events: generation %.2f ms, LSM %.2f ms)'
                             % (gen, lsm_ms)},
        'prices': [float(price[0])], 'clocks': clocks,
        'e2e': {'value': n * steps / (wall_ms * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': 50 * 8 * 2 + 148 * 6 * 8, 'd2h_bytes_per_step': 16},
        'gpu_launches': args.steps * (2 + 1 + 50 + 2),   # paths + column-sum reduce, init, passes, value sum + reduce
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE streaming pass at 8 M paths
                     # (ncu --set full, cache flushed before the launch: 192.1 MB read + 9.8 MB written
                     # against 256 MB algorithmic -- most of W's 64 MB store is still in L2 when the
                     # kernel ends): profiles/r1z_c5_lsm_step_fused.txt
                     'frac': achieved / hbm_peak, 'traffic': 201859840.0 if n == 8_000_000 else None,
                     'traffic_unit': 'bytes per streaming pass (ncu, 8M paths, cold L2)',
                     'note': 'LSM passes: 32 algorithmic bytes per path per exercise date (SURVEY 8d) '
                             '/ time between the device events around least_square_mc (initial payoff, 50 streaming '
                             'passes with fused solves, value sum); peak = MEASURED_PEAKS.json hbm_gbs'},
        'cpu_baseline': {'value': cv, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': '%d paths x %d steps + LSM, single process numpy oracle (%.1f s)' % (cn, csteps, cdt)},
    }
    emit(line)
  plan.close()
  if px is not None:
    px.close()
  if world > 1:
    dist.destroy_process_group()


_JSON_FD = None


def emit(line):
  """Writes the ONE JSON line of the run to the process's original stdout."""
  data = (json.dumps(line) + '\n').encode()
  if _JSON_FD is None:
    sys.stdout.write(data.decode())
    sys.stdout.flush()
  else:
    os.write(_JSON_FD, data)


def main():
  # stdout carries exactly one JSON line: anything else written to fd 1 during the
  # run (NCCL prints its version banner there) is sent to stderr instead
  global _JSON_FD
  sys.stdout.flush()
  _JSON_FD = os.dup(1)
  os.dup2(2, 1)
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
  ap.add_argument('--paths', type=int, default=None,
                  help='override the number of paths (parity / debugging only)')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  elif args.workload == 'c5':
    run_gpu_c5(args)
  else:
    run_gpu(args)


if __name__ == '__main__':
  main()
