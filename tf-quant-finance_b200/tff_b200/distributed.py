"""Sharding of a Monte-Carlo call over the GPUs of one box.

Paths are independent units: rank r of G owns the contiguous unit range
`shard_units(units, r, G)`.  The draws of a unit are a pure function of its
GLOBAL number (Philox element offset `p * S * dim`, Sobol index `skip + 1 + p`;
for antithetic types a unit is the pair (p, p + N/2)), so no rank needs
anything from another one until the reduction: one all-reduce of the
`[num_payoffs, 4]` unnormalised sums (NCCL over NVLink through
`torch.distributed`), plus, for Longstaff-Schwartz, the column sums and one
`K^2 + K` block per exercise date.
"""
import contextlib
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from tff_b200 import _lib


def world():
  """(rank, world_size) of the default process group, (0, 1) without one."""
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def shard_units(units, rank=None, world_size=None):
  """(offset, count) of the contiguous unit range owned by `rank`."""
  if rank is None or world_size is None:
    rank, world_size = world()
  units, rank, world_size = int(units), int(rank), int(world_size)
  per = (units + world_size - 1) // world_size
  lo = min(rank * per, units)
  hi = min((rank + 1) * per, units)
  return lo, hi - lo


def all_reduce_(tensor):
  """In-place sum over the default process group (no-op on a single rank)."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(tensor)
  return tensor


_SHARDED = None       # (peer_exchange or None) while a `sharded()` context is active


@contextlib.contextmanager
def sharded(peer_exchange=None):
  """Inside this context the pricing entry points of the package
  (`euler_sampling.price`, `GenericItoProcess.price`, `HestonModel.price`,
  `swaption_price`, `bond_option_price`, `cap_floor_price`) shard their paths over
  the ranks of the default process group and return GLOBAL prices on every rank:
  each rank simulates `shard_units(units)` and the `[num_payoffs, 4]` sums are
  added over NVLink peer memory (`peer_exchange`) or by one NCCL all-reduce.
  Collective: every rank must make the same calls."""
  global _SHARDED
  previous = _SHARDED
  _SHARDED = (peer_exchange,)
  try:
    yield
  finally:
    _SHARDED = previous


def price_sums(plan, payoffs):
  """`plan.price_sums(payoffs)` -- sharded over the ranks and globally reduced
  inside a `sharded()` context, the whole run on this GPU otherwise."""
  if _SHARDED is None or world()[1] == 1:
    plan.clear_peer_exchange()
    return plan.price_sums(list(payoffs))
  px = _SHARDED[0]
  lo, count = shard_units(plan.units)
  if px is not None and px.world > 1:
    plan.set_peer_exchange(px)
    return plan.price_sums(list(payoffs), lo, count)
  plan.clear_peer_exchange()
  return all_reduce_(plan.price_sums(list(payoffs), lo, count))


def price_sums_host(plan, payoffs):
  """`price_sums` read back to the host.  Sums that came through a peer exchange and
  are not finite are traced to the exchange's status words first: a time-out there
  raises instead of being returned as a NaN price."""
  s = price_sums(plan, payoffs).cpu().numpy()
  px = _SHARDED[0] if _SHARDED is not None else None
  if px is not None and px.world > 1 and not np.all(np.isfinite(s)):
    px.check('payoff sums')
  return s


def price_sharded(plan, payoffs, peer_exchange=None):
  """Monte-Carlo means of `payoffs` with the plan's units sharded over the
  ranks of the default process group.  Returns (mean, stderr) numpy arrays.
  With a `PeerExchange` the sums are added inside the reduction kernel over
  NVLink peer memory instead of by an NCCL all-reduce."""
  lo, count = shard_units(plan.units)
  if peer_exchange is not None and peer_exchange.world > 1:
    plan.set_peer_exchange(peer_exchange)
    sums = plan.price_sums(list(payoffs), lo, count)
  else:
    sums = plan.price_sums(list(payoffs), lo, count)
    all_reduce_(sums)
  s = sums.cpu().numpy()
  if peer_exchange is not None and peer_exchange.world > 1 and not np.all(np.isfinite(s)):
    peer_exchange.check('payoff sums')
  n = float(plan.num_samples)
  mean = s[:, 0] / n
  var = np.maximum(s[:, 1] / n - mean**2, 0.0)
  return mean, np.sqrt(var / n)


def paths_sharded(plan, record_slot, num_times, exp_transform=False):
  """This rank's rows of the path tensor (time-major view `[rows, k, dim]`) and
  the global index of its first unit."""
  lo, count = shard_units(plan.units)
  x = plan.paths(record_slot, num_times, lo, count, exp_transform)
  if plan.rng.antithetic and world()[1] > 1:
    # rows are [count units | their count partners]: row r >= count is global path
    # N/2 + lo + (r - count), not lo + r.  `least_square_mc(num_calibration_samples=)`
    # selects by global index lo + r and refuses tensors carrying this mark.
    x._tqf_antithetic_shard = True
  return x, lo


class PeerExchange:
  """Exchange buffers of the ranks of one box, mapped into every process (CUDA
  IPC over NVLink peer access).  `least_square_mc(..., peer_exchange=px)` then
  sums the per-date normal equations over the ranks inside the tail of its
  streaming kernel instead of calling NCCL once per exercise date.

  Collective: construct and `close()` on all ranks of `group` together, after
  `torch.cuda.set_device`.  At most 8 ranks, one per GPU (or several processes
  sharing a GPU, which time-slice)."""

  def __init__(self, group=None):
    self.group = group
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)
    if self.world > 8:
      raise ValueError('PeerExchange supports at most 8 ranks (one box)')
    lib = _lib.lib()
    self._own, self._opened, error = None, [], None
    handle_bytes = None
    try:
      _lib.require_cuda()
      nbytes = C.c_uint64()
      _lib.check(lib.tqf_lsm_peer_bytes(C.byref(nbytes)))
      own = C.c_void_p()
      handle = C.create_string_buffer(64)
      _lib.check(lib.tqf_peer_alloc(nbytes.value, C.byref(own), handle))
      self._own = own
      handle_bytes = handle.raw
    except Exception as e:  # pylint: disable=broad-except
      error = e
    # every rank takes part in both gathers whatever happened locally, so that a
    # failure on one rank (no IPC in this container, ...) fails ALL ranks together
    handles = [None] * self.world
    dist.all_gather_object(handles, handle_bytes, group=group)
    ptrs = []
    if error is None and all(h is not None for h in handles):
      try:
        for r, h in enumerate(handles):
          if r == self.rank:
            ptrs.append(self._own.value)
            continue
          p = C.c_void_p()
          _lib.check(lib.tqf_peer_open(C.create_string_buffer(h, 64), C.byref(p)))
          self._opened.append(p)
          ptrs.append(p.value)
      except Exception as e:  # pylint: disable=broad-except
        error = e
    elif error is None:
      error = RuntimeError('a peer could not allocate its exchange buffer')
    oks = [None] * self.world
    dist.all_gather_object(oks, error is None, group=group)
    if not all(oks):
      for p in self._opened:
        lib.tqf_peer_close(p)
      if self._own is not None:
        lib.tqf_peer_free(self._own)
      self._own, self._opened = None, []
      raise RuntimeError('PeerExchange could not be set up on every rank: {}'.format(
          error if error is not None else 'failure on another rank'))
    self.ptrs = (C.c_void_p * self.world)(*ptrs)
    self.epoch = 0          # exchanges performed so far (equal on all ranks)
    dist.barrier(group=group)

  def check(self, what='result'):
    """Raises if an in-kernel exchange on this rank's buffer gave up waiting for a
    peer (its sums were poisoned with NaN).  Called by the pricing entry points
    when a `what` that went through the exchange is not finite."""
    n, ep = C.c_uint64(), C.c_uint64()
    _lib.check(_lib.lib().tqf_peer_status(self._own, C.byref(n), C.byref(ep)))
    if n.value:
      raise RuntimeError(
          'non-finite {}: {} peer exchange(s) on rank {} timed out waiting for another rank '
          '(last at epoch {}); the ranks did not all reach the same exchange -- a failed or '
          'skipped call on one rank, or a rank more than ~10 s behind'.format(
              what, n.value, self.rank, ep.value))

  def close(self):
    if getattr(self, '_own', None) is None:
      return
    torch.cuda.synchronize()
    dist.barrier(group=self.group)       # nobody still writes into a peer
    lib = _lib.lib()
    for p in self._opened:
      lib.tqf_peer_close(p)
    dist.barrier(group=self.group)
    lib.tqf_peer_free(self._own)
    self._own, self._opened = None, []
