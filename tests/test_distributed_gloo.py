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
