"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the unit
partition, the global-index rule of the draws and the sum reduction.  The
kernels need a GPU, so each rank evaluates ITS unit range with the oracle --
exactly the (offset, count) arithmetic the engine hands to libtqf -- and the
all-reduced result must equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import models as omodels


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _heston(dtype=np.float64):
  return omodels.heston_closures(2.0, 0.04, 0.5, -0.7, dtype)


def _worker(rank, world_size, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world_size)
  try:
    from tff_b200 import distributed
    assert distributed.world() == (rank, world_size)
    n, steps = 1000, 8
    d, v = _heston()
    x0 = np.array([np.log(100.0), 0.04])
    res = {}
    # Sobol: a shard is the same call with skip advanced by its unit offset
    lo, cnt = distributed.shard_units(n)
    p = oeuler.sample(2, d, v, [1.0], num_time_steps=steps, num_samples=cnt,
                      initial_state=x0, random_type=odraws.RandomType.SOBOL,
                      skip=7 + lo, dtype=np.float64)
    sums = torch.tensor([np.maximum(np.exp(p[:, 0, 0]) - 100, 0).sum(), float(cnt)],
                        dtype=torch.float64)
    distributed.all_reduce_(sums)
    res['sobol'] = sums.numpy().copy()
    # Philox: element offset p * S * dim -> slice of the global draws tensor
    full = odraws.generate_mc_normal_draws(2, steps, n, odraws.RandomType.STATELESS,
                                           seed=[4, 2], dtype=np.float64)
    p = oeuler.sample(2, d, v, [1.0], num_time_steps=steps, initial_state=x0,
                      normal_draws=np.transpose(full[:, lo:lo + cnt], [1, 0, 2]),
                      dtype=np.float64)
    sums = torch.tensor([p[:, 0, 1].sum(), float(cnt)], dtype=torch.float64)
    distributed.all_reduce_(sums)
    res['philox'] = sums.numpy().copy()
    # antithetic: units are the first-half paths, each carries both partners
    half = n // 2
    lo, cnt = distributed.shard_units(half)
    anti = odraws.generate_mc_normal_draws(2, steps, n, odraws.RandomType.STATELESS_ANTITHETIC,
                                           seed=[4, 2], dtype=np.float64)
    rows = np.concatenate([np.arange(lo, lo + cnt), half + np.arange(lo, lo + cnt)])
    p = oeuler.sample(2, d, v, [1.0], num_time_steps=steps, initial_state=x0,
                      normal_draws=np.transpose(anti[:, rows], [1, 0, 2]), dtype=np.float64)
    sums = torch.tensor([p[:, 0, 0].sum(), float(2 * cnt)], dtype=torch.float64)
    distributed.all_reduce_(sums)
    res['anti'] = sums.numpy().copy()
    if rank == 0:
      np.save(out, res, allow_pickle=True)
  finally:
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
  out = str(tmp_path / 'res.npy')
  mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  res = np.load(out, allow_pickle=True).item()
  n, steps = 1000, 8
  d, v = _heston()
  x0 = np.array([np.log(100.0), 0.04])
  p = oeuler.sample(2, d, v, [1.0], num_time_steps=steps, num_samples=n, initial_state=x0,
                    random_type=odraws.RandomType.SOBOL, skip=7, dtype=np.float64)
  np.testing.assert_allclose(res['sobol'], [np.maximum(np.exp(p[:, 0, 0]) - 100, 0).sum(), n],
                             rtol=1e-13)
  p = oeuler.sample(2, d, v, [1.0], num_time_steps=steps, num_samples=n, initial_state=x0,
                    random_type=odraws.RandomType.STATELESS, seed=[4, 2], dtype=np.float64)
  np.testing.assert_allclose(res['philox'], [p[:, 0, 1].sum(), n], rtol=1e-13)
  p = oeuler.sample(2, d, v, [1.0], num_time_steps=steps, num_samples=n, initial_state=x0,
                    random_type=odraws.RandomType.STATELESS_ANTITHETIC, seed=[4, 2],
                    dtype=np.float64)
  np.testing.assert_allclose(res['anti'], [p[:, 0, 0].sum(), n], rtol=1e-13)


@pytest.mark.parametrize('units,world', [(10, 1), (10, 2), (10, 3), (7, 8), (0, 2),
                                         (10_000_000, 8), (25_000_000, 8)])
def test_shard_units_is_a_partition(units, world):
  from tff_b200 import distributed
  covered = 0
  prev_end = 0
  for r in range(world):
    lo, cnt = distributed.shard_units(units, r, world)
    assert lo == prev_end and cnt >= 0
    prev_end = lo + cnt
    covered += cnt
  assert covered == units and prev_end == units


def _peer_worker(rank, world_size, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world_size)
  try:
    from tff_b200 import distributed
    try:
      distributed.PeerExchange()
      result = 'constructed'
    except RuntimeError as e:
      result = 'RuntimeError: %s' % e
    with open(out % rank, 'w') as f:
      f.write(result)
  finally:
    dist.destroy_process_group()


def test_peer_exchange_fails_on_all_ranks_together_without_a_gpu(tmp_path):
  # no CUDA device here: the set-up must fail collectively (RuntimeError on every
  # rank, nobody left waiting in a barrier) so that callers can fall back to NCCL
  if torch.cuda.is_available():
    pytest.skip('needs a machine without a GPU')
  out = str(tmp_path / 'peer%d.txt')
  mp.spawn(_peer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  for r in range(2):
    assert open(out % r).read().startswith('RuntimeError: PeerExchange could not be set up')


# ----- sharded Longstaff-Schwartz: the three reductions of SURVEY 8e on two ranks
def _sharded_lsm(paths_local, exercise_times, strike, degree, df, all_reduce):
  """The algorithm the multi-GPU path executes (tff_b200 least_square_mc with
  `all_reduce`): (1) all-reduce of the column sums -> basis means, (2) per
  exercise date all-reduce of the masked normal equations, identical K x K solve
  on every rank, local update of the merged state W = cashflow + values,
  (3) all-reduce of the value sum.  numpy restatement, dim 1, one payoff."""
  x = paths_local[:, :, 0]                                       # [n_local, T]
  n_local, T = x.shape
  dfx = np.concatenate([[1.0], df])
  ratio = dfx[1:] / dfx[:-1]                                     # [T]: df[e + 1] / df[e]
  stats = torch.tensor(np.concatenate([x.sum(axis=0), [float(n_local)]]))
  all_reduce(stats)
  means = stats[:-1].numpy() / float(stats[-1])
  K = degree + 1
  w = np.maximum(strike - x[:, exercise_times[T - 1]], 0.0)      # cashflow at the last date
  for e in range(T - 1, 0, -1):
    t = exercise_times[e - 1]
    ev = np.maximum(strike - x[:, t], 0.0)
    phi = (x[:, t] - means[t])[:, None] ** np.arange(K)[None, :]   # [n, K]
    y = ratio[e] * w
    m = ev > 0
    sums = torch.tensor(np.concatenate([(phi[m].T @ phi[m]).reshape(-1), phi[m].T @ y[m]]))
    all_reduce(sums)
    lhs, rhs = sums[:K * K].numpy().reshape(K, K), sums[K * K:].numpy()
    beta = np.linalg.pinv(lhs, rcond=10 * K * np.finfo(np.float64).eps) @ rhs
    cont = np.maximum(phi @ beta, 0.0)
    w = np.where(ev > cont, ev, y)
  vs = torch.tensor([float((ratio[0] * w).sum()), float(n_local)])
  all_reduce(vs)
  return float(vs[0] / vs[1])


def _lsm_paths(n):
  rs = np.random.RandomState(11)
  times = np.linspace(0.0, 1.0, 9)
  z = rs.normal(size=(n, 8))
  logs = np.concatenate([np.zeros((n, 1)),
                         np.cumsum((0.06 - 0.2) * 0.125 + np.sqrt(0.4 * 0.125) * z, axis=1)], axis=1)
  return np.exp(logs)[..., None], np.exp(-0.06 * times)


def _lsm_worker(rank, world_size, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world_size)
  try:
    from tff_b200 import distributed
    paths, df = _lsm_paths(4000)
    lo, cnt = distributed.shard_units(paths.shape[0])
    price = _sharded_lsm(paths[lo:lo + cnt], np.arange(9), 1.1, 3, df,
                         lambda t: distributed.all_reduce_(t))
    np.save(out % rank, np.array([price]))
  finally:
    dist.destroy_process_group()


def test_two_rank_sharded_lsm_matches_single_process_oracle(tmp_path):
  from oracle import lsm as olsm
  out = str(tmp_path / 'lsm%d.npy')
  mp.spawn(_lsm_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  got = [float(np.load(out % r)[0]) for r in range(2)]
  assert got[0] == got[1]                    # every rank solved from the same reduced sums
  paths, df = _lsm_paths(4000)
  want = olsm.least_square_mc(paths, np.arange(9), olsm.make_basket_put_payoff([1.1]),
                              olsm.make_polynomial_basis(3), df, dtype=np.float64)
  # The normal equations are summed in a different order (ITM rows only, two halves) and the
  # cubic Gram matrix is ill-conditioned: beta moves by ~1e-9, and ONE path whose exercise
  # value sits that close to the fitted continuation value may flip (|ev - y| / N ~ 2e-8 here).
  # An algorithmic error (wrong discounting, wrong mask, wrong means) is of order 1e-2.
  np.testing.assert_allclose(got[0], want[0], rtol=1e-6)
  # and the sharded recursion on ONE rank is the same algorithm
  single = _sharded_lsm(paths, np.arange(9), 1.1, 3, df, lambda t: t)
  np.testing.assert_allclose(single, want[0], rtol=1e-6)
  np.testing.assert_allclose(got[0], single, rtol=1e-6)


# ---- the product's own sharded flow, the device plan replaced by the CPU stand-in -----------------
def _sharded_flow_worker(rank, world_size, port, out):
  """`distributed.sharded()` around the public pricing calls, `paths_sharded`: every line of host
  code the multi-GPU run executes, with `tests/cpu_plan.CpuPlan` in place of the device plan."""
  import sys
  import pytest as _pytest
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  import cpu_plan
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world_size)
  mp_ = _pytest.MonkeyPatch()
  try:
    cpu_plan.install(mp_)
    import tff_b200 as tff
    from tff_b200 import distributed, engine
    from tff_b200.models import euler_sampling
    mp_.setattr(euler_sampling, '_CALLS', type(euler_sampling._CALLS)())
    rt = tff.math.random.RandomType
    heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
    x0 = np.array([np.log(100.0), 0.04])
    payoffs = [engine.european_call(100.0, log_state=True), engine.up_and_out_call(100.0, 130.0, log_state=True)]
    res = {}
    for name, kw in (('sobol', dict(random_type=rt.SOBOL)),
                     ('anti', dict(random_type=rt.STATELESS_ANTITHETIC, seed=[4, 2]))):
      with distributed.sharded():
        res[name] = heston.price([1.0], payoffs, num_samples=1002, initial_state=x0, num_time_steps=8,
                                 return_stats=True, **kw)
    flat = lambda t: 0.01 * np.ones_like(np.asarray(t))
    with distributed.sharded():
      res['swaption'] = tff.models.hull_white.swaption_price(
          expiries=np.array([1.0]), floating_leg_start_times=None, floating_leg_end_times=None,
          floating_leg_daycount_fractions=None, fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0]]),
          fixed_leg_daycount_fractions=0.25 * np.ones((1, 4)), fixed_leg_coupon=0.011 * np.ones((1, 4)),
          reference_rate_fn=flat, notional=100., mean_reversion=0.03, volatility=0.02, num_samples=1000,
          time_step=0.1, seed=[4, 2], dtype=np.float64, use_analytic_pricing=False,
          random_type=rt.STATELESS_ANTITHETIC)
    # this rank's rows of a materialised antithetic run and the global index of its first unit
    gbm = tff.models.GeometricBrownianMotion(0.05, 0.3, dtype=np.float64)
    plans, record_slot, k, _ = euler_sampling._prepare(
        1, gbm.drift_fn(), gbm.volatility_fn(), [0.5, 1.0], None, 4, 10, [100.0], rt.STATELESS_ANTITHETIC, [1, 2],
        0, None, None, None, False, None, np.float64)
    rows, lo = distributed.paths_sharded(plans[0], record_slot, k)
    res['rows'], res['lo'] = rows.numpy().copy(), lo
    assert getattr(rows, '_tqf_antithetic_shard', False)
    np.save(os.path.join(out, 'flow_%d.npy' % rank), np.array([res], dtype=object), allow_pickle=True)
  finally:
    mp_.undo()
    dist.destroy_process_group()


def test_two_rank_sharded_product_flow_matches_one_rank(tmp_path, monkeypatch):
  port = _free_port()
  mp.spawn(_sharded_flow_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  r0, r1 = (np.load(os.path.join(str(tmp_path), 'flow_%d.npy' % r), allow_pickle=True)[0] for r in (0, 1))
  # the same calls on one rank
  import sys
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  import cpu_plan
  cpu_plan.install(monkeypatch)
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import euler_sampling
  monkeypatch.setattr(euler_sampling, '_CALLS', type(euler_sampling._CALLS)())
  rt = tff.math.random.RandomType
  heston = tff.models.HestonModel(mean_reversion=2.0, theta=0.04, volvol=0.5, rho=-0.7, dtype=np.float64)
  x0 = np.array([np.log(100.0), 0.04])
  payoffs = [engine.european_call(100.0, log_state=True), engine.up_and_out_call(100.0, 130.0, log_state=True)]
  for name, kw in (('sobol', dict(random_type=rt.SOBOL)), ('anti', dict(random_type=rt.STATELESS_ANTITHETIC, seed=[4, 2]))):
    one = heston.price([1.0], payoffs, num_samples=1002, initial_state=x0, num_time_steps=8, return_stats=True, **kw)
    for got in (r0[name], r1[name]):                   # GLOBAL prices on every rank
      for a, b in zip(got, one):
        np.testing.assert_allclose(a, b, rtol=1e-12)
  np.testing.assert_allclose(r0['swaption'], r1['swaption'], rtol=0, atol=0)
  from oracle import hull_white as ohw
  flat = lambda t: 0.01 * np.ones_like(np.asarray(t))
  want = ohw.swaption_price_mc(
      expiries=np.array([1.0]), fixed_leg_payment_times=np.array([[1.25, 1.5, 1.75, 2.0]]),
      fixed_leg_daycount_fractions=0.25 * np.ones((1, 4)), fixed_leg_coupon=0.011 * np.ones((1, 4)),
      reference_rate_fn=flat, notional=100., mean_reversion=0.03, volatility=0.02, num_samples=1000, time_step=0.1,
      seed=[4, 2], dtype=np.float64, random_type=odraws.RandomType.STATELESS_ANTITHETIC)
  np.testing.assert_allclose(r0['swaption'], want, rtol=1e-11)
  # materialised antithetic run of 10 paths = 5 units: rank 0 owns units 0..2, rank 1 units 3..4; rows are
  # [units | partners] and together they are the one-rank tensor
  gbm = tff.models.GeometricBrownianMotion(0.05, 0.3, dtype=np.float64)
  full = tff.models.euler_sampling.sample(1, gbm.drift_fn(), gbm.volatility_fn(), [0.5, 1.0], num_time_steps=4,
                                          num_samples=10, initial_state=[100.0], random_type=rt.STATELESS_ANTITHETIC,
                                          seed=[1, 2], dtype=np.float64).numpy()
  assert (r0['lo'], r1['lo']) == (0, 3) and r0['rows'].shape[0] == 6 and r1['rows'].shape[0] == 4
  np.testing.assert_array_equal(np.concatenate([r0['rows'][:3], r1['rows'][:2], r0['rows'][3:], r1['rows'][2:]]), full)
