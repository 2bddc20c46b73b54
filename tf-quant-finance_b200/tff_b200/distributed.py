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
import numpy as np
import torch
import torch.distributed as dist


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


def price_sharded(plan, payoffs):
  """Monte-Carlo means of `payoffs` with the plan's units sharded over the
  ranks of the default process group.  Returns (mean, stderr) numpy arrays."""
  lo, count = shard_units(plan.units)
  sums = plan.price_sums(list(payoffs), lo, count)
  all_reduce_(sums)
  s = sums.cpu().numpy()
  n = float(plan.num_samples)
  mean = s[:, 0] / n
  var = np.maximum(s[:, 1] / n - mean**2, 0.0)
  return mean, np.sqrt(var / n)


def paths_sharded(plan, record_slot, num_times, exp_transform=False):
  """This rank's rows of the path tensor (time-major view `[rows, k, dim]`) and
  the global index of its first unit."""
  lo, count = shard_units(plan.units)
  return plan.paths(record_slot, num_times, lo, count, exp_transform), lo
