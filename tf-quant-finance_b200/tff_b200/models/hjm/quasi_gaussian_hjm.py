"""Quasi-Gaussian HJM (`models/hjm/quasi_gaussian_hjm.py:33-545`) on the B200 path engine.

The reference simulates the state `[x, vec(y)]` (dimension F + F^2) with
`euler_sampling.sample`; the zero-volatility `y` rows still take part in the Wiener
process, so every Euler step consumes F + F^2 normals.  With a deterministic
volatility (a constant vector, or a callable that does not depend on the short
rate) `y` is the same for every path: its Euler recursion runs ONCE on the host and
enters the device model (TQF_MODEL_HJM) as the drift column `sum_j y_ij`; the kernel
carries the F factors and the running integral of the short rate that the
discount factors are made of, and consumes the same F + F^2 draws per step, so the
random stream is the reference's seed for seed.  A volatility that depends on the
short rate is refused (`NotImplementedError`: no CPU fallback).
"""
import ctypes as C

import numpy as np
import torch

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import engine
from tff_b200.models import utils
from tff_b200.models.hull_white import _exact


def _valid_sqrt_matrix(rho):
  """`_get_valid_sqrt_matrix` (528-545): Cholesky, or V sqrt(max(e, 1e-5)) when an
  eigenvalue is below 1e-5."""
  e, v = np.linalg.eigh(rho.astype(np.float64))
  if np.any(e < 1e-5):
    return v @ np.sqrt(np.diag(np.maximum(e, 1e-5)))
  return np.linalg.cholesky(rho.astype(np.float64))


class _HjmSpec(engine.ModelSpec):
  """TQF_MODEL_HJM with an explicit per-step table (columns in tqf.h)."""
  kind = _lib.MODEL_HJM

  def __init__(self, factors, draws_per_step, table):
    self.dim = factors + 1
    self.num_factors = int(draws_per_step)
    self.num_coef = 5 + 2 * factors + factors * factors
    self._table = table

  def coef_table(self, all_times, dtype):
    del all_times, dtype
    return self._table


class QuasiGaussianHJM:
  """`QuasiGaussianHJM(dim, mean_reversion, volatility, initial_discount_rate_fn,
  corr_matrix=None, validate_args=False, dtype=None, name=None)`; `dim` <= 2 factors
  (the quasi-Gaussian state of 3 factors would consume 12 draws per step)."""

  _DRAWS_PER_STEP_IS_STATE_DIM = True      # F + F^2 (state [x, vec(y)])
  _RIGHT_POINT_DISCOUNTING = True          # quasi_gaussian_hjm.py:486-490

  def __init__(self, dim, mean_reversion, volatility, initial_discount_rate_fn,
               corr_matrix=None, validate_args=False, dtype=None, name=None):
    self._name = name or 'quasi_gaussian_hjm_model'
    self._dtype = _tensor.np_dtype(dtype, np.float32)
    self._factors = int(dim)
    self._dim = self._factors + self._factors**2
    dt_ = self._dtype
    self._mean_reversion = _tensor.to_numpy(mean_reversion, dt_).reshape(-1)
    if self._mean_reversion.shape[0] != self._factors:
      raise NotImplementedError('batches of HJM models are not implemented by the B200 engine')
    self._initial_discount_rate_fn = initial_discount_rate_fn
    self._fwd, _ = _exact.forward_rate_fns(initial_discount_rate_fn, dt_)
    if callable(volatility):
      self._volatility = volatility
    else:
      vol = _tensor.to_numpy(volatility, dt_).reshape(self._factors)
      self._volatility = lambda t, r: vol
    rho = (np.eye(self._factors, dtype=dt_) if corr_matrix is None
           else _tensor.to_numpy(corr_matrix, dt_))
    if rho.shape != (self._factors, self._factors):
      raise NotImplementedError('batches of HJM models are not implemented by the B200 engine')
    self._rho = rho
    if validate_args:
      try:
        self._sqrt_rho = np.linalg.cholesky(rho.astype(np.float64)).astype(dt_)
      except np.linalg.LinAlgError:
        raise ValueError('The input correlation matrix is not positive semidefinite.')
    else:
      self._sqrt_rho = _valid_sqrt_matrix(rho).astype(dt_)

  # ------------------------------------------------------------ accessors --
  def dim(self):
    return self._dim

  def dtype(self):
    return self._dtype

  def name(self):
    return self._name

  def instant_forward_rate(self, t):
    return self._fwd(np.asarray(t, dtype=self._dtype))

  # ------------------------------------------------------------ host tables
  def _sigma(self, t):
    """sigma_i(t) [F] on the host; the callable must not depend on the short rate."""
    dt_ = self._dtype
    outs = []
    for r in (0.01, 0.37):
      val = None
      for mode in ('torch', 'numpy'):
        try:
          if mode == 'torch':
            tt = torch.tensor(float(t), dtype=_tensor.torch_dtype(dt_))
            rr = torch.full((1, 1), r, dtype=_tensor.torch_dtype(dt_))
            val = self._volatility(tt, rr)
            if isinstance(val, torch.Tensor):
              val = val.detach().cpu().numpy()
          else:
            val = self._volatility(dt_.type(t), np.full((1, 1), r, dtype=dt_))
          val = np.asarray(val, dtype=np.float64)
          break
        except Exception:  # pylint: disable=broad-except
          val = None
      if val is None:
        raise NotImplementedError(
            'could not evaluate the HJM volatility callable on the host (tried torch and numpy '
            'inputs)')
      outs.append(np.broadcast_to(val.reshape(-1, val.shape[-1]) if val.ndim else val,
                                  (1, self._factors)).reshape(self._factors))
    if not np.allclose(outs[0], outs[1], rtol=1e-12, atol=0):
      raise NotImplementedError(
          'The B200 HJM kernel runs deterministic volatilities sigma(t); this callable depends '
          'on the short rate (a genuinely quasi-Gaussian model). There is no CPU fallback.')
    return outs[0].astype(dt_)

  def _grids(self, times, time_step, num_time_steps):
    """The grid `_sample_paths` builds (451-463) and the one `euler_sampling.sample`
    builds from it (euler_sampling.py:232-283): returns (grid, idx of `times` in it,
    all_times of the Euler loop, its keep mask)."""
    dt_ = self._dtype
    if time_step is None and num_time_steps is None:
      raise ValueError('Either `time_step` or `num_time_steps` should be supplied.')
    if time_step is not None and num_time_steps is not None:
      raise ValueError('When `times_grid` is not supplied only one of either '
                       '`num_time_steps` or `time_step` should be defined but not both.')
    ts_internal = None if time_step is None else dt_.type(_tensor.to_numpy(time_step))
    if num_time_steps is not None:
      num_time_steps = int(num_time_steps)
      ts_internal = dt_.type(times[-1] / dt_.type(num_time_steps))
    grid, _, idx = utils.prepare_grid(times=times, time_step=ts_internal, dtype=dt_,
                                      num_time_steps=num_time_steps)
    ts2 = ts_internal if num_time_steps is None else dt_.type(grid[-1] / dt_.type(num_time_steps))
    all_times, keep_mask, _ = utils.prepare_grid(times=grid, time_step=ts2, dtype=dt_,
                                                 num_time_steps=num_time_steps)
    return grid, np.asarray(idx, dtype=np.int64), all_times, keep_mask

  def _y_columns(self, all_times, sigma):
    """y at the START of every Euler step, [S, F, F], and y after the last step: the
    Euler recursion of the reference's `y` state (drift 271-276, zero volatility)."""
    dt_ = self._dtype
    f = self._factors
    k = self._mean_reversion
    mr2 = (k[:, None] + k[None, :]).astype(dt_)
    steps = all_times.shape[0] - 1
    y = np.zeros((f, f), dtype=dt_)
    out = np.empty((steps + 1, f, f), dtype=dt_)
    for s in range(steps):
      out[s] = y
      dt = dt_.type(all_times[s + 1] - all_times[s])
      vol = sigma[s].reshape(f, 1)
      drift = (self._rho * (vol @ vol.T).astype(dt_) - mr2 * y).astype(dt_)
      y = (y + dt * drift).astype(dt_)
    out[steps] = y
    return out

  def _tables(self, all_times, integral_weights=None):
    """(coef table [S, NCOEF] float64, y at every Euler grid entry [S + 1, F, F]).

    `integral_weights` [S] replaces the model's own discounting rule: step s adds
    `w_s r(all_times[s + 1])` to the short-rate integral (the rule of
    `options_price_from_samples`, hjm/zero_coupon_bond_option_util.py:102-113)."""
    dt_ = self._dtype
    f = self._factors
    steps = all_times.shape[0] - 1
    dts = (all_times[1:] - all_times[:-1]).astype(dt_)
    sq = np.sqrt(dts).astype(dt_)
    sigma = np.stack([self._sigma(all_times[s + 1]) for s in range(steps)], 0) if steps else (
        np.zeros((0, f), dt_))
    y_entries = self._y_columns(all_times, sigma)
    a0 = self._drift_a0(all_times, y_entries)                       # [S, F]
    b = (self._sqrt_rho[None, :, :] * sigma[:, :, None]).astype(dt_)       # [S, F, F]
    f0 = np.asarray(self._fwd(all_times), dtype=dt_)
    if integral_weights is not None:
      w = np.asarray(integral_weights, dtype=dt_)
      c_l, c_r, c_f = np.zeros_like(dts), w, (f0[1:] * w).astype(dt_)
    elif self._RIGHT_POINT_DISCOUNTING:
      c_l, c_r, c_f = np.zeros_like(dts), dts, (f0[1:] * dts).astype(dt_)
    else:
      c_l, c_r, c_f = dts, np.zeros_like(dts), (f0[:-1] * dts).astype(dt_)
    cols = [dts[:, None], sq[:, None], a0, np.broadcast_to(self._mean_reversion, (steps, f)),
            b.reshape(steps, f * f), c_l[:, None], c_r[:, None], c_f[:, None]]
    return np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.float64), y_entries

  def _drift_a0(self, all_times, y_entries):
    del all_times
    return y_entries[:-1].sum(-1)                                   # sum_j y_ij at the step start

  def _draws_per_step(self):
    return self._dim if self._DRAWS_PER_STEP_IS_STATE_DIM else self._factors

  def _plan(self, times, time_step, num_time_steps, num_samples, random_type, seed, skip,
            integral_weights_fn=None):
    """The device plan over the Euler grid, which Euler entry each requested time is
    read at, the y tables at those entries, and the gather of duplicate times.
    `integral_weights_fn(all_times, entries)` (entries = Euler entry of every requested
    time) may replace the discounting rule, see `_tables`."""
    dt_ = self._dtype
    times = _tensor.to_numpy(times, dt_)
    if times.ndim != 1:
      raise ValueError('`times` should be a rank 1 Tensor. '
                       'Rank is {} instead.'.format(times.ndim))
    f = self._factors
    if (f, self._draws_per_step()) not in ((1, 1), (1, 2), (2, 2), (2, 6), (3, 3)):
      raise NotImplementedError(
          'The B200 HJM kernel covers 1-2 factors (quasi-Gaussian) and 1-3 factors (Gaussian).')
    grid, idx, all_times, keep_mask = self._grids(times, time_step, num_time_steps)
    num_steps, grid_slot = engine.record_plan(keep_mask, grid.shape[0])
    # requested time u lives in grid slot idx[u]; record each distinct slot once
    uniq, inverse = np.unique(idx, return_inverse=True)
    record_slot = np.full(num_steps + 1, -1, dtype=np.int32)
    entry_of = np.full(uniq.shape[0], -1, dtype=np.int64)
    for entry in range(num_steps + 1):
      g = grid_slot[entry]
      if g >= 0:
        pos = np.searchsorted(uniq, g)
        if pos < uniq.shape[0] and uniq[pos] == g:
          record_slot[entry] = pos
          entry_of[pos] = entry
    if np.any(entry_of < 0):
      raise ValueError('a requested time is not reached by the simulation grid')
    weights = None
    if integral_weights_fn is not None:
      weights = integral_weights_fn(all_times, entry_of[inverse])
    table, y_entries = self._tables(all_times, weights)
    spec = _HjmSpec(f, self._draws_per_step(), table[:num_steps])
    rng = engine.RngSpec(random_type, seed, skip, None)
    plan = engine.Plan(spec, all_times, num_steps, np.zeros(f + 1, dt_), rng, int(num_samples),
                       dt_)
    return plan, record_slot, entry_of, inverse, y_entries, times

  def _sample(self, times, time_step, num_time_steps, num_samples, random_type, seed, skip):
    """(state [N, k, F + 1] with the short-rate integral last, y [k, F, F], times)."""
    plan, record_slot, entry_of, inverse, y_entries, times = self._plan(
        times, time_step, num_time_steps, num_samples, random_type, seed, skip)
    try:
      x = plan.paths(record_slot, entry_of.shape[0])             # [N, unique times, F + 1]
    finally:
      plan.close()
    if inverse.shape[0] != entry_of.shape[0] or np.any(inverse != np.arange(inverse.shape[0])):
      x = x.index_select(1, torch.as_tensor(inverse, device=x.device))
    return x, self._y_at(times, y_entries[entry_of][inverse]), times

  def _y_at(self, times, y_simulated):
    del times
    return y_simulated

  # ------------------------------------------------------------ sampling ---
  def sample_paths(self, times, num_samples, time_step=None, num_time_steps=None,
                   random_type=None, seed=None, skip=0, name=None):
    """`(short rates [N, k], discount factors [N, k], x [N, k, F], y [N, k, F^2])`
    (`quasi_gaussian_hjm.py:291-363`), CUDA tensors (y is a broadcast view)."""
    del name
    state, y, times = self._sample(times, time_step, num_time_steps, num_samples, random_type,
                                   seed, skip)
    f = self._factors
    x = state[..., :f]
    f0 = torch.as_tensor(np.asarray(self._fwd(times), dtype=self._dtype), device=x.device)
    rates = x.sum(-1) + f0[None, :]
    df = torch.exp(-state[..., f])
    y_t = torch.as_tensor(y.reshape(times.shape[0], f * f), device=x.device, dtype=x.dtype)
    return rates, df, x, y_t[None].expand(x.shape[0], -1, -1)

  def _bond_tables(self, times, maturities, y):
    """A [m, k] and G [m, k, F] of P = A exp(-G . x) (`_bond_reconstitution`, 499-525):
    `times` [k], `maturities` [m, k], `y` [k, F, F]."""
    dt_ = self._dtype
    k = self._mean_reversion
    rate = lambda t: _exact.discount_rate(self._initial_discount_rate_fn, t, dt_)
    p0 = np.exp(-rate(maturities) * maturities) / np.exp(-rate(times) * times)[None, :]
    g = (1. - np.exp(-k * (maturities[..., None] - times[None, :, None]))) / k      # [m, k, F]
    term2 = np.einsum('mki,kij,mkj->mk', g, y, g)
    return (p0 * np.exp(-0.5 * term2)).astype(np.float64), g.astype(np.float64)

  def sample_discount_curve_paths(self, times, curve_times, num_samples, time_step=None,
                                  num_time_steps=None, random_type=None, seed=None, skip=0,
                                  name=None):
    """`(P(t, t + tau) [N, m, k], short rates [N, k], discount factors [N, k])`
    (`quasi_gaussian_hjm.py:365-449`); the bond curves are written by one CUDA kernel
    from the factor paths (`tqf_hjm_discount_curves`)."""
    del name
    dt_ = self._dtype
    curve_times = _tensor.to_numpy(curve_times, dt_)
    state, y, times = self._sample(times, time_step, num_time_steps, num_samples, random_type,
                                   seed, skip)
    f = self._factors
    a, g = self._bond_tables(times, times[None, :] + curve_times[:, None], y)
    dev = state.device
    a_dev = torch.as_tensor(np.ascontiguousarray(a), device=dev)
    g_dev = torch.as_tensor(np.ascontiguousarray(g), device=dev)
    n, m, k = int(state.shape[0]), int(curve_times.shape[0]), int(times.shape[0])
    out = _tensor.empty((n, m, k), dt_)
    _lib.check(_lib.lib().tqf_hjm_discount_curves(
        state.data_ptr(), state.stride(0), state.stride(1), state.stride(2), a_dev.data_ptr(),
        g_dev.data_ptr(), n, m, k, f, _tensor.tqf_dtype(dt_), out.data_ptr(),
        _tensor.current_stream_ptr()))
    x = state[..., :f]
    f0 = torch.as_tensor(np.asarray(self._fwd(times), dtype=dt_), device=dev)
    return out, x.sum(-1) + f0[None, :], torch.exp(-state[..., f])
