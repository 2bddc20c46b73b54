"""The per-date all-reduce of the Longstaff-Schwartz normal equations through
peer memory (`lsm.py:369-377` forms lhs / rhs per exercise date):
two processes (one GPU each when the box has two, otherwise time-slicing one
GPU -- CUDA IPC works either way) shard the paths, exchange the sums inside the
tail of the streaming kernel and must reproduce the single-process price."""
import os
import sys
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, SIGMA, N, T = 0.1, 1.0, 40_000, 13


def _setup(strikes=(1.1, 1.2)):
  sys.path.insert(0, os.path.join(ROOT, 'tf-quant-finance_b200'))
  import tff_b200 as tff
  from tff_b200 import engine
  from tff_b200.models import closures, utils
  times = np.linspace(0.0, 1.0, T)
  drift, vol = closures.affine_closures(R - SIGMA**2 / 2, 0.0, SIGMA)
  spec = closures.resolve_spec(drift, vol)
  all_times, mask, _ = utils.prepare_grid(times=times, time_step=np.float64(0.05), dtype=np.float64)
  steps, record_slot = engine.record_plan(mask, T)
  rng = engine.RngSpec(tff.math.random.RandomType.STATELESS, [4, 2], 0)
  plan = engine.Plan(spec, all_times, steps, np.array([0.0]), rng, N, np.float64)
  lsm = tff.models.longstaff_schwartz
  kw = dict(discount_factors=np.exp(-R * times), dtype=np.float64)
  put = lsm.make_basket_put_payoff(list(strikes), dtype=np.float64)
  return plan, record_slot, lsm, put, lsm.make_polynomial_basis(3), kw


def _worker(rank, world, port, out_path, strikes=(1.1, 1.2)):
  try:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank,
                            world_size=world)
    plan, record_slot, lsm, put, basis, kw = _setup(strikes)
    from tff_b200 import distributed
    lo, count = distributed.shard_units(plan.units, rank, world)
    paths, sums = plan.paths(record_slot, T, lo, count, exp_transform=True, column_sums=True)
    reduce_fn = lambda t: dist.all_reduce(t)
    px = distributed.PeerExchange()
    res = []
    for _ in range(2):   # twice on the same buffers: the epochs continue
      res.append(lsm.least_square_mc(paths, np.arange(T), put, basis, global_path_offset=lo,
                                     all_reduce=reduce_fn, column_sums=sums, peer_exchange=px, **kw))
    # one payoff: the persistent single-launch induction (T exchanges: T - 1 regressions
    # and the value sum); several payoffs: one launch per date (T - 1 exchanges)
    assert px.epoch == (2 * T if len(strikes) == 1 else 2 * (T - 1)), px.epoch
    # the NCCL-style route (one all-reduce per date) on the same shards
    res.append(lsm.least_square_mc(paths, np.arange(T), put, basis, global_path_offset=lo,
                                   all_reduce=reduce_fn, column_sums=sums, **kw))
    # fused pricing: the payoff sums of both ranks are added inside the reduction kernel
    from tff_b200 import engine
    payoffs = [engine.european_call(1.0, log_state=True), engine.european_put(1.2, log_state=True)]
    plan.set_peer_exchange(px)
    fused = [plan.price_sums(payoffs, lo, count).cpu().numpy() for _ in range(3)]
    np.save(out_path % (10 + rank), np.stack(fused))
    px.close()
    plan.close()
    np.save(out_path % rank, np.stack(res))
    dist.destroy_process_group()
  except Exception:  # pylint: disable=broad-except
    traceback.print_exc()
    os._exit(1)


@pytest.mark.parametrize('strikes', [(1.1, 1.2), (1.1,)], ids=['per_date_launches', 'persistent'])
def test_two_ranks_peer_exchange_matches_single_process(tmp_path, strikes):
  import torch
  import torch.multiprocessing as mp
  plan, record_slot, lsm, put, basis, kw = _setup(strikes)
  paths = plan.paths(record_slot, T, exp_transform=True)
  want = lsm.least_square_mc(paths, np.arange(T), put, basis, **kw)
  plan.close()
  torch.cuda.synchronize()
  ctx = mp.get_context('spawn')
  out = str(tmp_path / 'rank%d.npy')
  port = 29600 + os.getpid() % 300
  procs = [ctx.Process(target=_worker, args=(r, 2, port, out, strikes)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(240)
  hung = [p for p in procs if p.is_alive()]
  for p in hung:
    p.kill()
  assert not hung, 'peer exchange worker did not finish'
  assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
  # fused pricing: both ranks hold the global sums, equal to a single-process run
  from tff_b200 import engine
  plan, _, _, _, _, _ = _setup()
  payoffs = [engine.european_call(1.0, log_state=True), engine.european_put(1.2, log_state=True)]
  whole = plan.price_sums(payoffs).cpu().numpy()
  plan.close()
  fused = [np.load(out % (10 + r)) for r in range(2)]
  np.testing.assert_array_equal(fused[0], fused[1])
  for row in fused[0]:
    np.testing.assert_allclose(row[:, :2], whole[:, :2], rtol=1e-13)
    np.testing.assert_array_equal(row[:, 2], whole[:, 2])
  got = [np.load(out % r) for r in range(2)]
  # every rank solved from bit-identical sums -> identical prices on both ranks
  np.testing.assert_array_equal(got[0], got[1])
  np.testing.assert_array_equal(got[0][0], got[0][1])
  for row in got[0]:
    np.testing.assert_allclose(row, want, rtol=1e-9)
