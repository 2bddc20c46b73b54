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
