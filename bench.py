#!/usr/bin/env python
"""Benchmark of the fused Euler Monte-Carlo path engine (BASELINE.json metric:
Euler path-steps/sec, fp64).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--only]
                  [--impl reference]

One "step" = one pass of the hot path over the whole workload (all paths x all
Euler steps, normals generated in-kernel, payoffs reduced in-kernel).  The
headline workload is BASELINE.json configs[1] (C2): Heston Euler, 10M paths x
252 steps, float64, Sobol, European + up-and-out barrier call.  Unless `--only`
is given the same run then times every other BASELINE config with the same
event discipline and reports them in the `workloads` object of the ONE JSON
line: C1, C3, C4 (strict + clamped draws), C5 (generation / LSM split), the QE
scheme of C2 and the path-materialising mode recording every step.

Paths shard across ranks by disjoint Sobol index ranges / Philox counter
ranges (the total workload is the named config: "scaling": "strong"); the
payoff sums of the ranks are added inside the reduction kernel over NVLink
peer memory, the LSM normal equations inside the regression kernel.
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

# Algorithmic FP-pipe instructions per path-step (SURVEY.md 8(d);
# frozen in roofline.json -- SURVEY allows them to be tightened only downward).
ALGO_INSTR = {'c1': 12, 'c2': 79, 'c2_qe': 79, 'c3': 24, 'c4': 1408, 'c5': 17}
# (c1 / c3 / c5: tightened in round 2 to no more than the kernels EXECUTE -- ncu: 12.8 / 24.3 /
# 17.7 FP64 instructions per path-step with the table logarithm behind Box-Muller -- so that the
# fraction cannot exceed what the FP64 pipe does; round 1 froze 16 / 29 / 22)
# C4 is bound by the issue slots (and, next to them, the shared-memory pipe), not by one
# FP pipe: its contraction runs on the tensor cores.  A(C4) = 64 draws x 22 thread-instructions
# of ANY pipe (Sobol word 2.25, table inverse CDF 14, scale 1, TF32 split 3, update 1, staging
# 0.75), tightened down from SURVEY's 3616 FP32-pipe instructions (2080 of them were the
# mat-vec FMAs, 64 x 6 the polynomial inverse CDF); peak = issue rate = the FFMA rate.
ISSUE_BOUND = ('c4',)

WORKLOADS = {
    'c1': dict(name='C1 GBM call (log-space affine), 100k paths x 100 steps, fp64, PSEUDO_ANTITHETIC seed 42',
               paths=100_000, dtype='f64'),
    'c2': dict(name='C2 Heston Euler, 10M paths x 252 steps, fp64, Sobol, European + up-and-out call',
               paths=10_000_000, dtype='f64'),
    'c2_qe': dict(name='C2 Heston QE (HestonModel.sample_paths scheme), 10M paths x 252 steps, fp64, Sobol, '
                       'European + up-and-out call', paths=10_000_000, dtype='f64'),
    'c3': dict(name='C3 Hull-White 1F payer swaption (exact OU step + discount integral), 50M paths x 360 steps, '
                    'fp64, Philox stateless seed [4,2], through swaption_price',
               paths=50_000_000, dtype='f64'),
    'c4': dict(name='C4 correlated 64-asset GBM basket call, 20M paths x 252 steps, fp32, Sobol + Cholesky',
               paths=20_000_000, dtype='f32'),
    'c5': dict(name='C5 American put, Longstaff-Schwartz on log-GBM Euler paths, 8M paths x 50 exercise dates '
                    '(148 Euler steps, time_step 0.01), fp64, STATELESS_ANTITHETIC seed [4,2], cubic basis',
               paths=8_000_000, dtype='f64'),
    'materialise': dict(name='path-materialising mode: log-GBM Euler paths recorded at EVERY one of 64 steps, '
                             '8M paths, fp64, STATELESS_ANTITHETIC (4.1 GB written per pass)',
                        paths=8_000_000, dtype='f64'),
}
EXTRA_ORDER = ['c1', 'c3', 'c4', 'c5', 'c2_qe', 'materialise']

HESTON = dict(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7)
C3 = dict(expiries=np.array(1.0), fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
          fixed_leg_daycount_fractions=0.25 * np.ones(4), fixed_leg_coupon=0.011 * np.ones(4),
          mean_reversion=0.03, volatility=0.02, notional=100.0, seed=[4, 2],
          time_step=1.0 / 360, dtype=np.float64)


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
  """One worker of the CPU arm: the numpy oracle (op-for-op restatement of the
  reference's TF CPU path: precomputed draws tensor + one vectorised update per
  step) on `count` paths.  Returns (sum of payoffs, steps)."""
  name, skip, count = args
  from oracle import draws as odraws
  from oracle import euler as oeuler
  from oracle import models as omodels
  if name in ('c2', 'c2_qe'):
    x0 = np.array([np.log(100.0), 0.04])
    if name == 'c2':
      d, v = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, np.float64)
      paths, xmax, _ = oeuler.sample(
          2, d, v, [1.0], num_time_steps=252, num_samples=count, initial_state=x0,
          random_type=odraws.RandomType.SOBOL, skip=skip, dtype=np.float64, return_extrema=True)
    else:
      from oracle import heston_qe as oqe
      paths, xmax, _ = oqe.sample_paths(
          2.0, 0.04, 0.5, -0.7, [1.0], x0, num_samples=count, num_time_steps=252,
          random_type=odraws.RandomType.SOBOL, skip=skip, dtype=np.float64, return_extrema=True)
    call = np.maximum(np.exp(paths[:, -1, 0]) - 100.0, 0)
    # both payoffs of the config: European call and up-and-out call (barrier 130 on every grid point)
    return [float(call.sum()), float(np.where(np.exp(xmax) > 130.0, 0.0, call).sum())], 252
  if name == 'c1':
    r, sigma = 0.03, 0.1
    paths = oeuler.sample(
        1, lambda t, x: (r - sigma**2 / 2) + 0 * x,
        lambda t, x: sigma * np.ones(x.shape + (1,)), [1.0], time_step=0.01,
        num_samples=count, initial_state=np.array([np.log(700.0)]),
        random_type=odraws.RandomType.PSEUDO_ANTITHETIC, seed=42 + skip,
        dtype=np.float64)
    st = np.exp(paths[:, 0, 0])
    return [float(np.maximum(st - k, 0).sum()) for k in (600.0, 650.0, 680.0)], 100
  if name == 'c3':
    from oracle import hull_white as ohw
    _, payoff = ohw.swaption_price_mc(
        expiries=np.array(1.0), fixed_leg_payment_times=np.array([1.25, 1.5, 1.75, 2.0]),
        fixed_leg_daycount_fractions=0.25 * np.ones(4),
        fixed_leg_coupon=0.011 * np.ones(4), reference_rate_fn=lambda t: 0.01 + 0 * t,
        mean_reversion=0.03, volatility=0.02, notional=100., num_samples=count,
        random_type=odraws.RandomType.STATELESS, seed=[4, 2 + skip],
        time_step=1.0 / 360, dtype=np.float64, return_payoffs=True)
    return [float(payoff.sum())], 360
  if name == 'c4':
    dim = 64
    d, v = omodels.mvgbm_closures(np.full(dim, 0.03, np.float32),
                                  np.linspace(0.1, 0.4, dim).astype(np.float32),
                                  (0.3 + 0.7 * np.eye(dim)).astype(np.float32), np.float32)
    paths = oeuler.sample(dim, d, v, np.array([1.0], np.float32), num_time_steps=252,
                          num_samples=count, initial_state=100.0 * np.ones(dim, np.float32),
                          random_type=odraws.RandomType.SOBOL, skip=skip, dtype=np.float32)
    return [float(np.maximum(paths[:, 0, :].mean(axis=1) - 100.0, 0).sum())], 252
  if name in ('c5', 'materialise'):
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
    return [float(price[0]) * count], 148
  raise ValueError(name)


def cpu_run(name, sample_paths, procs):
  """Times the oracle on `sample_paths` paths split over `procs` worker
  processes.  Returns (path_steps_per_s, seconds, steps, paths)."""
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


CPU_SAMPLE_1CORE = {'c1': 100_000, 'c2': 32768, 'c2_qe': 32768, 'c3': 65536, 'c4': 1024, 'c5': 65536,
                    'materialise': 65536}


def run_reference(args):
  """`--impl reference`: the reference's CPU path.  TensorFlow cannot be
  installed in this image (no wheel, no network), so the oracle port -- the
  numpy restatement pinned by the reference's known-answer tests -- is timed
  on all host cores, pricing the same payoffs as the GPU arm."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  sample = {'c1': 100_000, 'c2': 8192 * cores, 'c2_qe': 8192 * cores, 'c3': 16384 * cores,
            'c4': 256 * cores, 'c5': 16384 * cores, 'materialise': 16384 * cores}[args.workload]
  for _ in range(args.warmup):
    cpu_run(args.workload, max(sample // 8, 2 * cores), cores)
  secs = []
  for _ in range(args.steps):
    _, dt, steps, n = cpu_run(args.workload, sample, cores)
    secs.append(dt)
  value = float(n * steps * len(secs) / np.sum(secs))
  w = WORKLOADS[args.workload]
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * float(np.mean(secs)), 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': w['dtype'],
      'data': 'synthetic',
      'config': {'workload': w['name'], 'sample_paths': n,
                 'note': 'bounded sample of the workload (same model, grid, generator and payoffs); '
                         'oracle port of the TF CPU path, rate-normalised'},
      'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                       'sample': '%d paths x %d steps per step, %d processes' % (n, steps, cores)},
      'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  emit(line)


# ------------------------------------------------------------- GPU side ----
class Ctx:
  """Process-wide state of the GPU arm: ranks, streams, the peer exchange."""

  def __init__(self, args):
    import torch
    import torch.distributed as dist
    self.torch, self.dist, self.args = torch, dist, args
    self.world = int(os.environ.get('WORLD_SIZE', '1'))
    self.rank = int(os.environ.get('RANK', '0'))
    self.local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(self.local)
    if self.world > 1:
      dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
    self.stream = torch.cuda.current_stream()
    self.flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
    self._align = torch.zeros(1, dtype=torch.float32, device='cuda')
    self.px = None
    if self.world > 1 and os.environ.get('TQF_PRICE_PEER_EXCHANGE', '1') != '0':
      from tff_b200 import distributed
      try:
        self.px = distributed.PeerExchange()      # fails on ALL ranks together or on none
      except RuntimeError as e:
        sys.stderr.write('peer exchange unavailable (%s): NCCL all-reduce of the sums\n' % e)
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # pylint: disable=broad-except
      pass
    self.hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    self.hbm_peak_kind = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6650'
    self.fma_peaks = None

  def shard(self, units):
    per = (units + self.world - 1) // self.world
    lo = min(self.rank * per, units)
    return lo, min((self.rank + 1) * per, units) - lo

  def sharding_note(self):
    if self.world == 1:
      return 'single GPU, no exchange'
    if self.px is not None:
      return ('disjoint path ranges per rank; sums added over NVLink peer memory inside the '
              'reduction kernels (no NCCL call in the step)')
    return 'disjoint path ranges per rank; NCCL all-reduce of the sums'

  def max_over_ranks(self, values):
    t = self.torch.tensor(list(values), dtype=self.torch.float64, device='cuda')
    if self.world > 1:
      self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]

  def barrier(self):
    self.torch.cuda.synchronize()
    if self.world > 1:
      self.dist.barrier()

  def timed(self, step_fn, steps, warmup, flush=True, sample_clocks=False):
    """W >= 3 untimed steps, then exactly `steps` steps, each bracketed by CUDA
    events on the launching stream (the L2 flush between iterations sits outside
    the events); barrier + synchronize on both sides; max over ranks.  Returns
    (ms_per_step, wall_s, clocks, last result)."""
    torch = self.torch
    out = None
    for _ in range(max(warmup, 3)):
      out = step_fn()
    # (the clock sampler spawns nvidia-smi: start it BEFORE the barrier, or rank 0 enters
    # the timed region ~0.5 ms after the others and every other rank waits for it inside
    # the first step's in-kernel exchange)
    sampler = ClockSampler(self.local) if (sample_clocks and self.rank == 0) else None
    if sampler:
      sampler.start()
      time.sleep(0.3)
    self.barrier()
    if self.world > 1:
      # device-side alignment: ranks leave the host barrier some 100 us apart; this
      # all-reduce is enqueued without blocking the host, so every GPU passes it at the
      # same moment with its first timed step already queued behind it
      self.dist.all_reduce(self._align)
    evs = []
    t0 = time.perf_counter()
    for _ in range(steps):
      if flush:
        self.flush_buf.zero_()                   # evict L2 between timed iterations
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(self.stream)
      out = step_fn()
      e1.record(self.stream)
      evs.append((e0, e1))
    self.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    ms = sum(a.elapsed_time(b) for a, b in evs)
    if os.environ.get('TQF_BENCH_DEBUG'):
      sys.stderr.write('rank %d: per-step ms %s\n' % (
          self.rank, ' '.join('%.3f' % a.elapsed_time(b) for a, b in evs)))
    ms, = self.max_over_ranks([ms])
    return ms / steps, wall, clocks, out

  def timed_wall(self, fn, steps):
    """End-to-end timing of `fn` (host buffers in, host result out): wall clock
    over `steps` calls between synchronising barriers, max over ranks."""
    for _ in range(3):                       # W >= 3 untimed calls
      fn()
    self.barrier()
    t0 = time.perf_counter()
    per_call = []
    for _ in range(steps):
      c0 = time.perf_counter()
      fn()                                   # returns host values: the call is synchronous
      per_call.append(time.perf_counter() - c0)
    self.barrier()
    dt, = self.max_over_ranks([time.perf_counter() - t0])
    if max(per_call) > 2.0 * min(per_call) + 1e-3:
      # one call far off the others (allocator growth, a host hiccup): say so instead of
      # folding it silently into the mean
      sys.stderr.write('e2e: uneven calls on rank %d: %s ms\n' % (
          self.rank, ' '.join('%.3f' % (1e3 * t) for t in per_call)))
    return dt / steps

  def peaks(self):
    if self.fma_peaks is None:
      from tff_b200 import engine
      self.fma_peaks = engine.measure_fma_peaks()
    return self.fma_peaks


def _fp_roofline(ctx, name, per_gpu_rate, note=''):
  dfma, ffma = ctx.peaks()
  fp32 = WORKLOADS[name]['dtype'] == 'f32'
  algo = ALGO_INSTR[name]
  achieved = per_gpu_rate * algo / 1e9
  peak = (ffma if fp32 else dfma) / 1e9
  traffic = None
  try:
    traffic = json.load(open(os.path.join(ROOT, 'roofline.json'))).get(name, {}).get('ncu', {}).get('dram_bytes')
  except Exception:  # pylint: disable=broad-except
    pass
  if name in ISSUE_BOUND:
    return {'bound': 'issue', 'achieved': achieved, 'peak': peak,
            'unit': 'G thread-instr/s (all pipes)', 'frac': achieved / peak,
            'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu capture named in roofline.json)',
            'note': 'achieved = path-steps/s/GPU x %d algorithmic thread-instructions per path-step '
                    '(roofline.json: draws, TF32 split and update; the [128 x 72] x [72 x 64] '
                    'contraction per tile and step runs on the tensor pipe: tcgen05.mma.kind::tf32, '
                    'ncu sm__pipe_tensor_cycles_active 20 %%); peak = issue slots = 128 lanes/clk/SM, '
                    'measured live as the FFMA rate by tqf_measure_fp64_peak; no per-path HBM '
                    'traffic%s' % (algo, note)}
  return {'bound': 'fp32' if fp32 else 'fp64', 'achieved': achieved, 'peak': peak,
          'unit': 'G %s-pipe instr/s' % ('FP32' if fp32 else 'FP64'), 'frac': achieved / peak,
          'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu capture named in roofline.json)',
          'note': 'achieved = path-steps/s/GPU x %d algorithmic %s instr per path-step (roofline.json); '
                  'peak = %s issue rate measured live by tqf_measure_fp64_peak (MEASURED_PEAKS.json has '
                  'no FP64/FP32 entry); the fused kernel has no per-path HBM traffic%s'
                  % (algo, 'FP32' if fp32 else 'FP64', 'FFMA' if fp32 else 'DFMA', note)}


class FusedWorkload:
  """A fused-price workload: `plan.price_sums(payoffs)` on this rank's shard."""

  def __init__(self, ctx, name, paths):
    import tff_b200 as tff
    from tff_b200 import engine
    from tff_b200.models import closures
    from tff_b200.models import utils
    self.ctx, self.name = ctx, name
    self.n = n = int(paths or WORKLOADS[name]['paths'])
    self.tff, self.engine = tff, engine
    rt = tff.math.random.RandomType
    self.grid_pricing = None
    self.clamped = False
    if name in ('c2', 'c2_qe'):
      self.model = tff.models.HestonModel(dtype=np.float64, **HESTON)
      self.x0 = np.array([np.log(100.0), 0.04])
      self.payoffs = [engine.european_call(100.0, log_state=True),
                      engine.up_and_out_call(100.0, 130.0, log_state=True)]
      self.kw = dict(num_samples=n, initial_state=self.x0, random_type=rt.SOBOL, num_time_steps=252)
      if name == 'c2':
        spec = closures.resolve_spec(self.model.drift_fn(), self.model.volatility_fn())
        all_times, mask, _ = utils.prepare_grid(
            times=np.array([1.0]), time_step=np.float64(1.0 / 252), num_time_steps=252, dtype=np.float64)
        steps, _ = engine.record_plan(mask, 1)
        self.plan = engine.Plan(spec, all_times, steps, self.x0, engine.RngSpec(rt.SOBOL, None, 0), n,
                                np.float64)
      else:
        from tff_b200.models.heston import qe
        self.plan, _, _ = qe._plan(self.model, [1.0], self.x0, n, rt.SOBOL, None, None, 0, 1e-6, 252,
                                   None, None)
      self.steps = self.plan.num_steps
      self.public = lambda i=0: self.model.price(
          [1.0], self.payoffs, scheme='qe' if name == 'c2_qe' else 'euler',
          **dict(self.kw, initial_state=self.x0 + np.array([1e-12 * i, 0.0])))
      self.h2d = self.steps * self.plan.spec.num_coef * 8 + 16 + self.plan.num_steps_total * 2 * 32 * 4
    elif name == 'c1':
      r, sigma = 0.03, 0.1
      d, v = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
      self.process = tff.models.GenericItoProcess(1, d, v, dtype=np.float64)
      self.x0 = np.array([np.log(700.0)])
      self.payoffs = [engine.european_call(k, log_state=True, scale=np.exp(-r))
                      for k in (600.0, 650.0, 680.0)]
      all_times, mask, _ = utils.prepare_grid(times=np.array([1.0]), time_step=np.float64(0.01),
                                              dtype=np.float64)
      steps, _ = engine.record_plan(mask, 1)
      self.plan = engine.Plan(closures.resolve_spec(d, v), all_times, steps, self.x0,
                              engine.RngSpec(rt.PSEUDO_ANTITHETIC, 42, 0), n, np.float64)
      self.steps = steps
      self.public = lambda i=0: self.process.price(
          [1.0], self.payoffs, num_samples=n, initial_state=self.x0 + 1e-12 * i,
          random_type=rt.PSEUDO_ANTITHETIC, seed=42, time_step=0.01)
      self.h2d = steps * 6 * 8 + 8
    elif name == 'c3':
      # swaption_test.py:30-44, 81-125 scaled up, through the public pricer
      from tff_b200.models.hull_white import swaption as swp
      self.c3kw = dict(C3, floating_leg_start_times=np.array([1.0, 1.25, 1.5, 1.75]),
                       floating_leg_end_times=np.array([1.25, 1.5, 1.75, 2.0]),
                       floating_leg_daycount_fractions=0.25 * np.ones(4),
                       reference_rate_fn=lambda t: 0.01 + 0 * t, use_analytic_pricing=False,
                       num_samples=n, random_type=rt.STATELESS)
      self.grid_pricing = swp.swaption_price(_plan_only=True, **self.c3kw)
      self.plan, self.payoffs = self.grid_pricing.plan, self.grid_pricing.payoffs
      self.steps = self.plan.num_steps
      self.public = lambda i=0: swp.swaption_price(
          **dict(self.c3kw, mean_reversion=self.c3kw['mean_reversion'] * (1.0 + 1e-12 * i)))
      self.h2d = self.steps * 5 * 8 + 16
    elif name == 'c4':
      dim = 64
      self.mv = tff.models.MultivariateGeometricBrownianMotion(
          dim, means=np.full(dim, 0.03, np.float32),
          volatilities=np.linspace(0.1, 0.4, dim).astype(np.float32),
          corr_matrix=(0.3 + 0.7 * np.eye(dim)).astype(np.float32), dtype=np.float32)
      spec = closures.resolve_spec(self.mv.drift_fn(), self.mv.volatility_fn())
      times = np.array([1.0], dtype=np.float32)
      all_times, mask, _ = utils.prepare_grid(
          times=times, time_step=np.float32(1.0) / np.float32(252), num_time_steps=252, dtype=np.float32)
      self.x0 = 100.0 * np.ones(dim, dtype=np.float32)
      self.payoffs = [engine.european_call(100.0, component=-1)]
      steps, _ = engine.record_plan(mask, 1)
      self.plan = engine.Plan(spec, all_times, steps, self.x0, engine.RngSpec(rt.SOBOL, None, 0), n,
                              np.float32)
      self.steps = steps
      def public(i=0):
        x0 = self.x0.copy()
        x0[0] += np.float32(1e-4) * i          # one float32 ulp steps of 100 are 7.6e-6
        return tff.models.euler_sampling.price(
            dim, self.mv.drift_fn(), self.mv.volatility_fn(), times, self.payoffs, num_time_steps=252,
            num_samples=n, initial_state=x0, random_type=rt.SOBOL, dtype=np.float32)
      self.public = public
      self.h2d = steps * 2 * 4 + dim * dim * 4 + 3 * dim * 4 + self.plan.num_steps_total * dim * 32 * 4
    else:
      raise ValueError(name)
    self.lo, self.count = ctx.shard(self.plan.units)
    if ctx.px is not None:
      self.plan.set_peer_exchange(ctx.px)

  def step(self):
    sums = self.plan.price_sums(self.payoffs, self.lo, self.count)
    if self.ctx.world > 1 and self.ctx.px is None:
      self.ctx.dist.all_reduce(sums)
    return sums

  def e2e(self, fresh=True):
    """One call of the public pricer.  `fresh`: with a parameter set no earlier call had
    (an input perturbed in its last digits), so that nothing is re-used from the plan
    cache: the tables are rebuilt on the host and uploaded inside the timed region."""
    from tff_b200 import distributed
    self._calls = getattr(self, '_calls', 0) + 1
    i = self._calls if fresh else 0
    if self.ctx.world > 1:
      with distributed.sharded(self.ctx.px):
        return self.public(i)
    return self.public(i)

  def result(self, ms_per_step, sums, e2e_s):
    ctx, n = self.ctx, self.n
    s = sums.cpu().numpy()
    rows = self.count * (2 if self.plan.rng.antithetic else 1)
    per_gpu_rate = rows * self.steps / (ms_per_step * 1e-3)
    note = ''
    if self.name == 'c3':
      note = ('; C3 is bound by the dispatch port, not by the FP64 pipe: the Philox rounds TensorFlow\'s '
              'stream prescribes are integer instructions beside the FP64 ones (roofline.json)')
    res = {
        'workload': WORKLOADS[self.name]['name'] if ctx.args.paths is None
                    else WORKLOADS[self.name]['name'] + ' [paths=%d]' % n,
        'value': n * self.steps / (ms_per_step * 1e-3), 'unit': UNIT, 'ms_per_step': ms_per_step,
        'paths': n, 'euler_steps': self.steps, 'dtype': WORKLOADS[self.name]['dtype'],
        'prices': (s[:, 0] / n).tolist(), 'non_finite': s[:, 2].tolist(),
        'gpu_launches_per_step': 2,
        'roofline': _fp_roofline(ctx, self.name, per_gpu_rate, note),
    }
    if self.name == 'c3':
      # what actually bounds C3: an FP64 instruction holds the dispatch port for two cycles,
      # every other instruction for one (tools/microbench/dfma_bench.cu)
      _, ffma = ctx.peaks()
      slots = 2 * ALGO_INSTR['c3'] + 20 + 10
      res['dispatch_roofline'] = {
          'bound': 'dispatch port', 'achieved': per_gpu_rate * slots / 1e9, 'peak': ffma / 1e9,
          'unit': 'G thread dispatch slots/s', 'frac': per_gpu_rate * slots / ffma,
          'note': 'algorithmic dispatch slots per path-step: 2 x %d FP64 + 20 (ten Philox rounds per four '
                  'words: 2 IMAD.WIDE + 2 LOP3 each, two normals per call) + 10 (counter, uint -> double '
                  'conversions, quadrant selects); peak = one slot per lane and clock = the FFMA rate '
                  'measured live; ncu: smsp__issue_active 63.8 %%, 69.6 executed instructions per '
                  'path-step of which 24.3 FP64 (profiles/r2f_c3.txt)' % ALGO_INSTR['c3']}
    if e2e_s is not None:
      res['e2e'] = {'value': n * self.steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': self.h2d,
                    'd2h_bytes_per_step': len(self.payoffs) * 4 * 8,
                    'api': {'c2': 'HestonModel.price', 'c2_qe': "HestonModel.price(scheme='qe')",
                            'c1': 'GenericItoProcess.price', 'c3': 'hull_white.swaption_price',
                            'c4': 'euler_sampling.price'}[self.name]}
    return res

  def close(self):
    if self.grid_pricing is not None:
      self.grid_pricing.close()
    else:
      self.plan.close()


def run_fused(ctx, name, steps, warmup, with_e2e=True, sample_clocks=False):
  w = FusedWorkload(ctx, name, ctx.args.paths)
  try:
    ms, wall, clocks, sums = ctx.timed(w.step, steps, warmup, sample_clocks=sample_clocks)
    e2e_s = ctx.timed_wall(w.e2e, max(1, min(steps, 3))) if with_e2e else None
    rep_s = ctx.timed_wall(lambda: w.e2e(fresh=False), max(1, min(steps, 3))) if with_e2e else None
    res = w.result(ms, sums, e2e_s)
    if rep_s is not None:
      res['e2e']['inputs'] = ('a parameter set no earlier call had, every call: tables rebuilt on the '
                              'host and uploaded inside the timed region (no plan-cache hit)')
      res['e2e']['repeated_call_value'] = w.n * w.steps / rep_s
      res['e2e']['repeated_call_note'] = ('the same call with identical arguments: device tables and, '
                                          'for constant-parameter models, the bound call are re-used')
    if name == 'c4':
      # beyond 2^24 points the float32 Sobol uniform can be exactly 1.0
      # (the reference's own erfinv returns +inf there); the strict run above DROPS those
      # paths and counts them, the clamped run maps u = 1.0 to the largest float32 below 1
      res['strict'] = {'price': res['prices'][0], 'non_finite_paths': res['non_finite'][0]}
      try:
        w.plan.set_sobol_clamp(True)
        ms_c, _, _, sums_c = ctx.timed(w.step, max(1, min(steps, 2)), 1)
        sc = sums_c.cpu().numpy()
        res['clamped'] = {'price': float(sc[0, 0] / w.n), 'non_finite_paths': float(sc[0, 2]),
                          'ms_per_step': ms_c}
      except (AttributeError, RuntimeError) as e:
        res['clamped'] = {'unavailable': str(e)}
      finally:
        try:
          w.plan.set_sobol_clamp(False)
        except (AttributeError, RuntimeError):
          pass
      # A/B: the same plan on the mma.sync (legacy HMMA) kernel the tcgen05 kernel replaced
      res['kernel'] = 'mvgbm_tc5_kernel (tcgen05.mma.kind::tf32, A and accumulator in tensor memory)'
      old_env = os.environ.get('TQF_MVGBM_TC5')
      os.environ['TQF_MVGBM_TC5'] = '0'
      try:
        ms_l, _, _, sums_l = ctx.timed(w.step, max(1, min(steps, 2)), 1)
        sl = sums_l.cpu().numpy()
        res['mma_sync_kernel'] = {'ms_per_step': ms_l, 'price': float(sl[0, 0] / w.n),
                                  'non_finite_paths': float(sl[0, 2])}
      finally:
        if old_env is None:
          del os.environ['TQF_MVGBM_TC5']
        else:
          os.environ['TQF_MVGBM_TC5'] = old_env
    res['wall_s_timed_region'] = wall
    if clocks is not None:
      res['clocks'] = clocks
    return res
  finally:
    w.close()


def run_c5(ctx, steps, warmup, sample_clocks=False):
  """C5: materialise 8M x 50 log-GBM Euler paths (time-major) and run the
  Longstaff-Schwartz backward induction on them.  One step = generation + LSM."""
  torch, dist = ctx.torch, ctx.dist
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  from tff_b200.models import utils
  lsm = tff.models.longstaff_schwartz
  w = WORKLOADS['c5']
  n = int(ctx.args.paths or w['paths'])
  r, sigma = 0.1, 1.0
  times = np.linspace(0.0, 1.0, 50)
  drift, vol = closures.affine_closures(r - sigma**2 / 2, 0.0, sigma)
  spec = closures.resolve_spec(drift, vol)
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.01), dtype=np.float64)
  nsteps, record_slot = engine.record_plan(mask, 50)
  rng = engine.RngSpec(tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0)
  plan = engine.Plan(spec, all_times, nsteps, np.array([0.0]), rng, n, np.float64)
  lo, count = ctx.shard(plan.units)
  df = np.exp(-r * times)
  put = lsm.make_basket_put_payoff([1.1], dtype=np.float64)
  basis = lsm.make_polynomial_basis(3)
  reduce_fn = (lambda t: dist.all_reduce(t)) if ctx.world > 1 else None
  split = {'gen': [], 'lsm': []}
  buf = {'paths': None}

  def one_step():
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(ctx.stream)
    # [rows, 50, 1] time-major view of exp(log-price), exponentiated on store; the kernel
    # that writes the paths also sums every column (the LSM basis means)
    paths, csums = plan.paths(record_slot, 50, lo, count, exp_transform=True, column_sums=True)
    e1.record(ctx.stream)
    price = lsm.least_square_mc(paths, np.arange(50), put, basis, discount_factors=df,
                                dtype=np.float64, global_path_offset=2 * lo, all_reduce=reduce_fn,
                                column_sums=csums, peer_exchange=ctx.px)
    e2.record(ctx.stream)
    torch.cuda.synchronize()
    split['gen'].append(e0.elapsed_time(e1))
    split['lsm'].append(e1.elapsed_time(e2))
    return price

  calls = {'n': 0}

  def fresh_call():
    # end to end with a parameter set no earlier call had: a new plan (tables rebuilt on the
    # host and uploaded), paths, backward induction, price read back
    calls['n'] += 1
    fresh = engine.Plan(spec, all_times, nsteps, np.array([1e-12 * calls['n']]), rng, n, np.float64)
    try:
      paths, csums = fresh.paths(record_slot, 50, lo, count, exp_transform=True, column_sums=True)
      price = lsm.least_square_mc(paths, np.arange(50), put, basis, discount_factors=df,
                                  dtype=np.float64, global_path_offset=2 * lo, all_reduce=reduce_fn,
                                  column_sums=csums, peer_exchange=ctx.px)
      return float(price[0])
    finally:
      fresh.close()

  try:
    ms, wall, clocks, price = ctx.timed(one_step, steps, warmup, flush=False, sample_clocks=sample_clocks)
    e2e_s = ctx.timed_wall(fresh_call, max(1, min(steps, 3)))
    gen = float(np.mean(split['gen'][-steps:]))
    lsm_ms = float(np.mean(split['lsm'][-steps:]))
    gen, lsm_ms = ctx.max_over_ranks([gen, lsm_ms])
    rows = 2 * count
    lsm_bytes = 49.0 * rows * 32.0
    achieved = lsm_bytes / (lsm_ms * 1e-3) / 1e9
    traffic = None
    try:
      traffic = json.load(open(os.path.join(ROOT, 'roofline.json'))).get('c5', {}).get('ncu', {}).get('dram_bytes')
    except Exception:  # pylint: disable=broad-except
      pass
    res = {
        'workload': w['name'] if ctx.args.paths is None else w['name'] + ' [paths=%d]' % n,
        'value': n * nsteps / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'paths': n,
        'euler_steps': nsteps, 'exercise_dates': 50, 'dtype': 'f64', 'prices': [float(price[0])],
        'generation_ms': gen, 'lsm_ms': lsm_ms,
        'l2': 'inputs (3.2 GB of paths) exceed L2',
        'e2e': {'value': n * nsteps / e2e_s, 'unit': UNIT,
                'h2d_bytes_per_step': 50 * 8 * 2 + nsteps * 6 * 8, 'd2h_bytes_per_step': 16,
                'api': 'engine.Plan + Plan.paths + longstaff_schwartz.least_square_mc',
                'inputs': 'a parameter set no earlier call had, every call: a new plan (tables rebuilt on '
                          'the host and uploaded), paths, backward induction, price read back',
                'repeated_call_value': n * nsteps / (wall / steps),
                'repeated_call_note': 'the timed region itself: the same plan re-used'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': ctx.hbm_peak, 'unit': 'GB/s',
                     'frac': achieved / ctx.hbm_peak, 'traffic': traffic,
                     'traffic_unit': 'bytes per date of the backward induction (ncu capture named in roofline.json)',
                     'note': 'LSM part: 32 algorithmic bytes per path per exercise date (SURVEY 8d: two path '
                             'columns, read + write of the merged state) x 49 dates / device time between the '
                             'events around least_square_mc; peak = HBM copy bandwidth, ' + ctx.hbm_peak_kind},
        'generation_roofline': _fp_roofline(ctx, 'c5', rows * nsteps / (gen * 1e-3),
                                            '; generation part only (paths stored at 50 of 148 steps)'),
        'wall_s_timed_region': wall,
    }
    if clocks is not None:
      res['clocks'] = clocks
    return res
  finally:
    plan.close()


def run_materialise(ctx, steps, warmup):
  """Path-materialising mode with cheap normals: every step recorded, time-major
  [k][dim][N] buffer, 8 bytes written per path and step."""
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures
  from tff_b200.models import utils
  w = WORKLOADS['materialise']
  n = int(ctx.args.paths or w['paths'])
  k = 64
  times = np.linspace(1.0 / k, 1.0, k)
  drift, vol = closures.affine_closures(0.03 - 0.02, 0.0, 0.2)
  spec = closures.resolve_spec(drift, vol)
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(1.0 / k), dtype=np.float64)
  nsteps, record_slot = engine.record_plan(mask, k)
  rng = engine.RngSpec(tff.math.random.RandomType.STATELESS_ANTITHETIC, [4, 2], 0)
  plan = engine.Plan(spec, all_times, nsteps, np.array([0.0]), rng, n, np.float64)
  lo, count = ctx.shard(plan.units)
  try:
    ms, wall, _, out = ctx.timed(lambda: plan.paths(record_slot, k, lo, count), steps, warmup, flush=False)
    rows = 2 * count
    nbytes = float(rows) * k * 8
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {
        'workload': w['name'], 'value': n * nsteps / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
        'paths': n, 'euler_steps': nsteps, 'dtype': 'f64', 'l2': 'output (4.1 GB) exceeds L2',
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': ctx.hbm_peak, 'unit': 'GB/s',
                     'frac': achieved / ctx.hbm_peak, 'traffic': nbytes,
                     'traffic_unit': 'bytes written per launch (algorithmic: 8 B per path and recorded step)',
                     'note': 'bytes stored / device time; the generator (Philox + Box-Muller in FP64) is what '
                             'bounds this mode, see generation_roofline of c5 and DESIGN.md 4.1'},
        'wall_s_timed_region': wall,
    }
  finally:
    plan.close()
    del out


def run_workload(ctx, name, steps, warmup, headline=False):
  if name == 'c5':
    return run_c5(ctx, steps, warmup, sample_clocks=headline)
  if name == 'materialise':
    return run_materialise(ctx, steps, warmup)
  return run_fused(ctx, name, steps, warmup, sample_clocks=headline)


def run_gpu(args):
  ctx = Ctx(args)
  t_start = time.perf_counter()
  head = run_workload(ctx, args.workload, args.steps, args.warmup, headline=True)
  extras = {}
  if not args.only and args.workload == 'c2' and args.paths is None:
    for name in EXTRA_ORDER:
      k = args.steps if name != 'c4' else max(2, min(args.steps, 3))
      try:
        extras[name] = run_workload(ctx, name, k, args.warmup)
        extras[name]['steps'] = k
      except Exception as e:  # pylint: disable=broad-except
        # collective state may be inconsistent after a failure on one rank: stop here
        extras[name] = {'error': repr(e)}
        break
  if ctx.rank == 0:
    cores = 1
    cv, cdt, csteps, cn = cpu_run(args.workload, CPU_SAMPLE_1CORE[args.workload], cores)
    w = WORKLOADS[args.workload]
    launches = {'c5': 2 + 1 + 2}.get(args.workload, head.get('gpu_launches_per_step', 2))
    line = {
        'metric': METRIC, 'value': head['value'], 'unit': UNIT, 'n_gpus': ctx.world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': head['ms_per_step'],
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': w['dtype'], 'data': 'synthetic',
        'config': {'workload': head['workload'], 'paths': head['paths'], 'euler_steps': head['euler_steps'],
                   'sharding': ctx.sharding_note(),
                   'l2': head.get('l2', 'flushed (256 MiB memset) between timed iterations; the kernel reads '
                                        '<100 KB of tables')},
        'prices': head.get('prices'), 'clocks': head.get('clocks'),
        'e2e': head.get('e2e') or {'value': head['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                                   'd2h_bytes_per_step': 0, 'api': 'engine.Plan.paths (device-resident output)'},
        'gpu_launches': launches * args.steps,
        'wall_s_timed_region': head['wall_s_timed_region'], 'roofline': head['roofline'],
        'cpu_baseline': {'value': cv, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d paths x %d steps, single process numpy oracle (%.1f s)' % (cn, csteps, cdt)},
    }
    for key in ('non_finite', 'strict', 'clamped', 'generation_ms', 'lsm_ms', 'generation_roofline'):
      if key in head:
        line[key] = head[key]
    if extras:
      line['workloads'] = extras
      line['workloads_note'] = ('the other BASELINE.json configs and modes, timed in this same run with the same '
                                'event discipline (warm-up >= 3, CUDA events on the launching stream, max over '
                                'ranks); each entry carries its own roofline')
    line['driver_run_s'] = time.perf_counter() - t_start
    emit(line)
  if ctx.px is not None:
    ctx.px.close()
  if ctx.world > 1:
    ctx.dist.destroy_process_group()


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
  ap.add_argument('--only', action='store_true',
                  help='time the headline workload only (skip the `workloads` object)')
  ap.add_argument('--paths', type=int, default=None,
                  help='override the number of paths (parity / debugging only)')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_gpu(args)


if __name__ == '__main__':
  main()
