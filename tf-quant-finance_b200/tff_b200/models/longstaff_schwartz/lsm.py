"""Longstaff-Schwartz least-squares Monte Carlo on the device
(`tf_quant_finance/models/longstaff_schwartz/lsm.py`).

Per exercise date the reference runs payoff_fn, basis_fn, two matmuls, a
pseudo-inverse and several `tf.where` passes over `[N]` tensors.  Here one
streaming kernel per date updates the merged state W = cashflow + values and
accumulates the normal equations of the next (earlier) date; the K x K
pseudo-inverse (numpy, rcond 10 K eps as `tf.linalg.pinv`) runs on the reduced
sums, after the all-reduce when paths are sharded over GPUs.
"""
import ctypes as C

import os

import numpy as np
import torch

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200.models.longstaff_schwartz import payoff_utils


class PolynomialBasis:
  """All monomials prod_j (x_j - mean_j)^{k_j}, 0 <= k_j <= degree
  (`lsm.py:50-125`).  Callable on the host like the reference's `basis`."""

  def __init__(self, degree):
    self.degree = int(degree)

  def exponents(self, dim):
    grid = np.arange(self.degree + 1)
    mesh = np.meshgrid(*(dim * [grid]))                     # 'xy' like tf.meshgrid
    return np.stack(mesh, -1).reshape(-1, dim).astype(np.int32)

  def __call__(self, sample_paths, time_index):
    x = sample_paths if isinstance(sample_paths, torch.Tensor) else torch.as_tensor(
        np.asarray(sample_paths))
    if x.dim() == 3:
      x = x.unsqueeze(0)
    sl = x[:, :, int(time_index):int(time_index) + 1, :]
    c = sl - sl.mean(dim=1, keepdim=True)
    e = torch.as_tensor(self.exponents(x.shape[-1]), dtype=x.dtype, device=x.device)
    return torch.prod(c**e, dim=-1).permute(0, 2, 1)


def make_polynomial_basis(degree):
  """Produces a callable from samples to a polynomial basis (`lsm.py:50-125`)."""
  return PolynomialBasis(degree)


def _unpack_sums(sums, K, packed):
  """[B, NS] reduced sums -> (lhs [B, K, K], rhs [B, K])."""
  B = sums.shape[0]
  if not packed:
    return sums[:, :K * K].reshape(B, K, K), sums[:, K * K:K * K + K]
  lhs = np.zeros((B, K, K))
  idx = 0
  for i in range(6):
    for j in range(i, 6):
      if i < K and j < K:
        lhs[:, i, j] = sums[:, idx]
        lhs[:, j, i] = sums[:, idx]
      idx += 1
  return lhs, sums[:, 21:21 + K]


def least_square_mc(sample_paths, exercise_times, payoff_fn, basis_fn,
                    discount_factors=None, num_calibration_samples=None,
                    dtype=None, name=None, *, global_path_offset=0,
                    all_reduce=None, column_sums=None, peer_exchange=None,
                    diagnostics=None):
  """Values Amercian style options using the LSM algorithm (`lsm.py:128-295`).

  Args are those of the reference; `exercise_times` are indices into the time
  axis of `sample_paths` (`[num_samples, num_times, dim]` or
  `[batch_size, num_samples, num_times, dim]`, a CUDA tensor -- e.g. the
  time-major view returned by the path engine -- or anything DLPack/numpy).
  `payoff_fn` / `basis_fn` must come from `make_basket_put_payoff` /
  `make_polynomial_basis` (arbitrary Python callables cannot run in a kernel).

  Extension for sharded paths: `global_path_offset` is the global index of the
  first local path and `all_reduce(tensor)` sums a device tensor in place over
  the ranks (e.g. `torch.distributed.all_reduce`).  `column_sums`: optional
  float64 device tensor `[num_times, dim]` with the sums of `sample_paths` over
  the (local) samples, as `engine.Plan.paths(..., column_sums=True)` returns
  them; the pass that forms the basis-centring means is then skipped.
  `peer_exchange`: a `tff_b200.distributed.PeerExchange` of the ranks that
  `all_reduce` spans; the per-date normal equations are then summed over the
  ranks inside the streaming kernel (peer memory over NVLink) and `all_reduce`
  is only called for the column sums and the final value sum.

  `diagnostics`: optional dict; when the persistent single-launch route runs it
  receives `sums` (`[T - 1, 27]`, row j = the reduced normal equations of
  exercise index `T - 1 - j` in the packed layout of `tqf_lsm_sums_layout`),
  `beta` (`[T - 1, 6]`), `w` (the final merged state `cashflow + values` of the
  local paths, a device tensor) and `route`.

  Returns a numpy array `[batch_size]`.
  """
  del name
  tabulated = isinstance(payoff_fn, payoff_utils.TabulatedPayoff)
  if not (tabulated or isinstance(payoff_fn, payoff_utils.BasketPutPayoff)) or not isinstance(
      basis_fn, PolynomialBasis):
    raise NotImplementedError(
        'The B200 LSM passes evaluate the payoff and the basis inside CUDA '
        'kernels: use make_basket_put_payoff(strikes) or '
        'make_tabulated_payoff(values), and make_polynomial_basis(degree). '
        'There is no CPU fallback.')
  if num_calibration_samples and getattr(sample_paths, '_tqf_antithetic_shard', False):
    raise ValueError(
        'num_calibration_samples with a shard of antithetic paths: the rows of the shard are '
        '[units | partners], so "the first num_calibration_samples paths" would depend on the '
        'number of ranks.  Draw the paths with a non-antithetic random type, or calibrate on '
        'all paths.')
  if isinstance(sample_paths, torch.Tensor) or hasattr(sample_paths, '__dlpack__'):
    x = _tensor.from_dlpack(sample_paths)
    if dtype is not None:
      x = x.to(_tensor.torch_dtype(_tensor.np_dtype(dtype)))
  else:
    x = torch.as_tensor(_tensor.to_numpy(sample_paths, None if dtype is None
                                         else _tensor.np_dtype(dtype)))
  x = x.to(_tensor.device())
  dt = _tensor.np_dtype(x.dtype)
  batched = x.dim() == 4
  if x.dim() not in (3, 4):
    raise ValueError('sample_paths must have rank 3 or 4')
  n_local, dim = int(x.shape[-3]), int(x.shape[-1])
  ex_times = np.asarray(_tensor.to_numpy(exercise_times)).astype(np.int64).reshape(-1)
  T = ex_times.shape[0]
  num_times = int(x.shape[-2])
  if T == 0 or ex_times.min() < 0 or ex_times.max() >= num_times:
    # the kernels index the time axis with these: the reference's tf.gather raises too
    raise ValueError('exercise_times must be indices into the time axis of sample_paths '
                     '(0 <= index < {}); got {}'.format(num_times, ex_times.tolist()))
  ev_tab = None
  if tabulated:
    vals = payoff_fn.values.to(device=x.device, dtype=x.dtype)        # [times, N, B]
    if int(vals.shape[1]) != n_local:
      raise ValueError('tabulated payoff does not match the number of sample paths')
    B = int(vals.shape[2])
    if batched and int(x.shape[0]) not in (1, B):
      raise ValueError('payoff batch does not match the batch of sample paths')
    strikes = np.zeros(B, dtype=np.float64)
    # [T, B, N]: one contiguous column per (exercise date, payoff)
    ev_tab = vals[torch.as_tensor(ex_times, device=vals.device)].permute(0, 2, 1).contiguous()
  else:
    strikes = payoff_fn.strikes.astype(np.float64)
    B = int(x.shape[0]) if batched else strikes.shape[0]
    if batched and strikes.shape[0] not in (1, B):
      raise ValueError('strikes batch does not match the batch of sample paths')
    strikes = np.ascontiguousarray(np.broadcast_to(strikes, (B,)))

  # discount factors: [T+1, B] with a leading 1 (lsm.py:239-256)
  ratio_path = None
  per_path = (isinstance(discount_factors, torch.Tensor) and discount_factors.dim() == 3
              and int(discount_factors.shape[0]) != 1)
  if per_path:
    # rank 3, one discount curve per sample: [N, 1, T] (lsm.py:208-256)
    dfp = discount_factors.to(device=x.device, dtype=x.dtype)
    if int(dfp.shape[0]) != n_local or int(dfp.shape[1]) != 1 or int(dfp.shape[2]) != T:
      raise NotImplementedError(
          'per-sample discount factors must have shape [num_samples, 1, num_exercise_times]')
    dfp = torch.cat([torch.ones_like(dfp[:, :, :1]), dfp], dim=-1)[:, 0, :]      # [N, T+1]
    ratio_path = (dfp[:, 1:] / dfp[:, :-1]).transpose(0, 1).contiguous()       # [T, N]
    ratio = np.ones((T, B), dtype=np.float64)
  else:
    if discount_factors is None:
      df = np.ones((1, 1, T), dtype=dt)
    else:
      df = _tensor.to_numpy(discount_factors, dt)
    if df.ndim == 0:
      df = df.reshape(1, 1, 1)
    if df.ndim == 1:
      df = df.reshape(1, 1, -1)
    if df.ndim != 3:
      raise NotImplementedError('discount_factors must have rank 0, 1 or 3')
    if df.shape[0] != 1:
      raise NotImplementedError(
          'per-sample discount factors must be a CUDA tensor [num_samples, 1, T]')
    df = np.concatenate([np.ones(df.shape[:2] + (1,), dtype=dt), df], axis=-1)
    df = np.broadcast_to(np.transpose(df, [2, 0, 1])[:, 0, :], (df.shape[-1], B))
    ratio = (df[1:] / df[:-1]).astype(dt).astype(np.float64)        # [T, B]

  exps = np.ascontiguousarray(basis_fn.exponents(dim), dtype=np.int32)
  K = exps.shape[0]
  stream = _tensor.current_stream_ptr()
  d = _lib.LsmDesc()
  d.dtype, d.dim, d.batch, d.basis_size = _tensor.tqf_dtype(dt), dim, B, K
  d.exponents, d.strikes = exps.ctypes.data, strikes.ctypes.data
  d.num_paths, d.path_offset = n_local, int(global_path_offset)
  d.num_calibration_samples = int(num_calibration_samples or 0)
  ex_i32 = np.ascontiguousarray(ex_times, dtype=np.int32)
  if ev_tab is not None or ratio_path is not None:
    if K > 6:
      raise NotImplementedError(
          'tabulated payoffs / per-sample discount factors need a basis of at most 6 functions')
    d.exercise_time_indices, d.num_exercise_times = ex_i32.ctypes.data, T
    d.exercise_values_dev = 0 if ev_tab is None else ev_tab.data_ptr()
    d.path_ratio_dev = 0 if ratio_path is None else ratio_path.data_ptr()
  d.paths_dev = x.data_ptr()
  st = x.stride()
  if batched:
    d.stride_batch, d.stride_path, d.stride_time, d.stride_dim = st
  else:
    d.stride_batch = 0
    d.stride_path, d.stride_time, d.stride_dim = st
  handle = C.c_void_p()
  lib = _lib.lib()
  _lib.require_cuda()
  # workspaces from torch's caching allocator (cudaMalloc / cudaFree per call
  # would synchronise the device)
  need = C.c_uint64()
  _lib.check(lib.tqf_lsm_workspace(C.byref(d), T, C.byref(need)))
  w_buf = torch.empty((B, max(n_local, 1)), dtype=x.dtype, device=x.device)
  part_buf = torch.empty((int(need.value),), dtype=torch.float64, device=x.device)
  d.w_dev, d.partials_dev, d.partials_doubles = w_buf.data_ptr(), part_buf.data_ptr(), need.value
  _lib.check(lib.tqf_lsm_create(C.byref(d), C.byref(handle)))
  try:
    ns, packed = C.c_int(), C.c_int()
    _lib.check(lib.tqf_lsm_sums_layout(handle, C.byref(ns), C.byref(packed)))
    dev = x.device
    # column means over ALL samples (lsm.py:110-111)
    tidx = np.ascontiguousarray(ex_times, dtype=np.int32)
    if column_sums is not None and not batched:
      cs = column_sums.to(device=dev, dtype=torch.float64)
      if cs.dim() != 2 or int(cs.shape[0]) != int(x.shape[-2]) or int(cs.shape[1]) != dim:
        raise ValueError('column_sums must have shape [num_times, dim] of sample_paths')
      colsum = cs[torch.as_tensor(ex_times, device=dev)].unsqueeze(0).expand(B, T, dim).contiguous()
    else:
      colsum = torch.zeros((B, T, dim), dtype=torch.float64, device=dev)
      _lib.check(lib.tqf_lsm_column_sums(handle, tidx.ctypes.data, T,
                                         colsum.data_ptr(), stream))
    count = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    if all_reduce is not None:
      all_reduce(colsum)
      all_reduce(count)
    # means in the working dtype (the reference reduces in dtype), as doubles;
    # everything below stays on the device: no host round trip per date.
    tdt = _tensor.torch_dtype(dt)
    means = (colsum / count).to(tdt).to(torch.float64).contiguous()     # [B, T, dim]
    ratio_dev = torch.as_tensor(np.ascontiguousarray(ratio), device=dev)    # [T, B]
    mean_stride = T * dim

    def mean_ptr(e):         # exercise index e uses time slot e - 1
      return means.data_ptr() + (e - 1) * dim * 8

    def ratio_ptr(e):
      return ratio_dev.data_ptr() + e * B * 8

    sums = torch.zeros((B, ns.value), dtype=torch.float64, device=dev)
    rcond = 10 * K * float(np.finfo(dt).eps)
    device_solve = bool(packed.value)
    multi = all_reduce is not None
    peers_possible = (multi and peer_exchange is not None and B <= 16
                      and os.environ.get('TQF_LSM_PEER_EXCHANGE', '1') != '0')

    def all_ranks_agree(ok):
      if not multi:
        return bool(ok)
      flag = torch.tensor([float(bool(ok))], dtype=torch.float64, device=dev)
      all_reduce(flag)
      return float(flag.item()) == float(peer_exchange.world)

    # Route 1: the whole backward induction in ONE persistent cooperative launch
    # (single asset, one payoff, K <= 6, contiguous time-major paths); several GPUs
    # exchange the per-date sums over NVLink peer memory inside it.
    persistent = False
    if (device_solve and B == 1 and not batched and (not multi or peers_possible)
        and os.environ.get('TQF_LSM_PERSISTENT', '1') != '0'):
      ok = C.c_int()
      _lib.check(lib.tqf_lsm_persistent_eligible(handle, C.byref(ok)))
      persistent = all_ranks_agree(ok.value) if multi else bool(ok.value)
    if persistent:
      if multi:
        _lib.check(lib.tqf_lsm_set_peer_exchange(handle, peer_exchange.rank, peer_exchange.world,
                                                 peer_exchange.ptrs, peer_exchange.epoch))
      history = None
      if diagnostics is not None:
        history = torch.zeros((max(T - 1, 1), 33), dtype=torch.float64, device=dev)
      vs = torch.zeros((B, 2), dtype=torch.float64, device=dev)
      beta_dev = torch.zeros((6,), dtype=torch.float64, device=dev)
      _lib.check(lib.tqf_lsm_run_persistent(
          handle, ex_i32.ctypes.data, T, means.data_ptr(), mean_stride, ratio_dev.data_ptr(),
          rcond, int(num_calibration_samples or 0), vs.data_ptr(), beta_dev.data_ptr(),
          None if history is None else history.data_ptr(), stream))
      if multi:
        ep = C.c_uint64()
        _lib.check(lib.tqf_lsm_peer_epoch(handle, C.byref(ep)))
        peer_exchange.epoch = int(ep.value)
      vs = vs.cpu().numpy()
      st = C.c_uint64()
      _lib.check(lib.tqf_lsm_status(handle, C.byref(st)))
      if st.value != 0:
        raise RuntimeError(
            'the persistent Longstaff-Schwartz kernel reported status {} (1: a CTA did not reach '
            'the grid barrier in time, 2 / 4: a tile never arrived, 3: a peer rank did not take '
            'part in the exchange within its time-out)'.format(st.value))
      if diagnostics is not None:
        h = history.cpu().numpy()
        diagnostics.update(route='persistent', sums=h[:T - 1, :27], beta=h[:T - 1, 27:],
                           w=w_buf[0])
      return (ratio[0] * vs[:, 0] / vs[:, 1]).astype(dt)

    _lib.check(lib.tqf_lsm_init(handle, int(ex_times[T - 1]), stream))
    beta_dev = torch.zeros((B, K), dtype=torch.float64, device=dev)
    e = T - 1
    # single rank + device solve: the per-CTA partials are reduced inside the
    # solve kernel (one launch less per date)
    fused = device_solve and all_reduce is None and B <= 128
    use_peers = False
    if device_solve and peers_possible:
      # every rank must take the same route: the fused pass has to apply everywhere
      ok = C.c_int()
      _lib.check(lib.tqf_lsm_fused_eligible(handle, C.byref(ok)))
      use_peers = all_ranks_agree(ok.value)
      fused = use_peers
    sums_arg = None if fused else sums.data_ptr()
    fused_set = fused and (use_peers or os.environ.get('TQF_LSM_FUSED_SOLVE', '1') != '0')
    if fused_set:
      # the last CTA of each streaming pass reduces and solves (no solve launch)
      ticket = torch.zeros((1,), dtype=torch.int32, device=dev)
      _lib.check(lib.tqf_lsm_set_fused_solve(handle, rcond, sums.data_ptr(),
                                             beta_dev.data_ptr(), ticket.data_ptr()))
      if use_peers:
        _lib.check(lib.tqf_lsm_set_peer_exchange(handle, peer_exchange.rank, peer_exchange.world,
                                                 peer_exchange.ptrs, peer_exchange.epoch))
    native_loop = False
    if fused_set:
      ok = C.c_int()
      _lib.check(lib.tqf_lsm_fused_eligible(handle, C.byref(ok)))
      native_loop = bool(ok.value)
    if native_loop:
      # every pass reduces and solves in its own tail: the whole backward
      # induction is launched back to back from native code
      _lib.check(lib.tqf_lsm_run_fused(handle, ex_i32.ctypes.data, T, means.data_ptr(),
                                       mean_stride, ratio_dev.data_ptr(), beta_dev.data_ptr(),
                                       stream))
      e = 0
    if e > 0:
      _lib.check(lib.tqf_lsm_step(handle, 0, 0, None, None, None, 1,
                                  int(ex_times[e - 1]), mean_ptr(e), ratio_ptr(e),
                                  mean_stride, sums_arg, stream))
    while e > 0:
      if all_reduce is not None and not use_peers:
        all_reduce(sums)
      if device_solve:
        _lib.check(lib.tqf_lsm_solve(handle, sums.data_ptr(), 1 if fused else 0, rcond,
                                     beta_dev.data_ptr(), stream))
      else:
        # large bases (K > 6): the K x K pseudo-inverse runs on the host
        lhs, rhs = _unpack_sums(sums.cpu().numpy(), K, False)
        beta = np.stack([np.linalg.pinv(lhs[b].astype(dt), rcond=rcond).astype(np.float64)
                         @ rhs[b].astype(dt).astype(np.float64) for b in range(B)])
        beta_dev.copy_(torch.as_tensor(np.ascontiguousarray(beta.astype(dt).astype(np.float64))))
      do_acc = 1 if e - 1 > 0 else 0
      _lib.check(lib.tqf_lsm_step(
          handle, 1, int(ex_times[e - 1]), mean_ptr(e), beta_dev.data_ptr(), ratio_ptr(e),
          do_acc, int(ex_times[e - 2]) if do_acc else 0,
          mean_ptr(e - 1) if do_acc else None, ratio_ptr(e - 1) if do_acc else None,
          mean_stride, sums_arg, stream))
      e -= 1
    vs = torch.zeros((B, 2), dtype=torch.float64, device=dev)
    _lib.check(lib.tqf_lsm_value_sum(handle, int(num_calibration_samples or 0),
                                     vs.data_ptr(), stream))
    if all_reduce is not None:
      all_reduce(vs)
    if use_peers:
      ep = C.c_uint64()
      _lib.check(lib.tqf_lsm_peer_epoch(handle, C.byref(ep)))
      peer_exchange.epoch = int(ep.value)
    vs = vs.cpu().numpy()
    if diagnostics is not None:
      diagnostics.update(route='per-date launches', w=w_buf[0])
    if use_peers and not np.all(np.isfinite(vs)):
      peer_exchange.check('Longstaff-Schwartz value sums')
      raise RuntimeError('non-finite Longstaff-Schwartz value sums after a peer exchange: a peer '
                         'rank did not take part within the time-out (rank skew or a failure on '
                         'another rank), or the paths hold non-finite values')
  finally:
    lib.tqf_lsm_destroy(handle)
  # per-sample discounting: the value sum already carries df[1] / df[0] per path
  return (ratio[0] * vs[:, 0] / vs[:, 1]).astype(dt)
