"""Oracle (test infrastructure): Longstaff-Schwartz least-squares Monte Carlo.

Restates `models/longstaff_schwartz/lsm.py`:
  * `make_polynomial_basis` 50-125, `least_square_mc` 128-295,
    `_apply_discount` 304-325, `_expected_exercise_fn` 328-383,
    `_updated_cashflows_and_values` 386-400, `_lsm_loop_body` 403-436
and `models/longstaff_schwartz/payoff_utils.py:27-97`
(`make_basket_put_payoff`).  `tf.linalg.pinv` (SVD, rcond = 10 max(rows, cols)
eps) is `numpy.linalg.pinv` with the same rcond.
"""
import numpy as np


def make_polynomial_basis(degree):
  """`lsm.py:50-125`: paths [B?, N, T, dim], time index -> [B, (deg+1)^dim, N]."""
  def basis(sample_paths, time_index):
    x = np.asarray(sample_paths)
    if x.ndim == 3:
      x = x[None]
    dim = x.shape[-1]
    sl = x[:, :, time_index:time_index + 1, :]               # [B, N, 1, dim]
    centered = sl - sl.mean(axis=1, keepdims=True)
    grid = np.arange(degree + 1, dtype=x.dtype)
    mesh = np.meshgrid(*(dim * [grid]))                       # 'xy' like tf.meshgrid
    grid = np.stack(mesh, -1).reshape(-1, dim)                # [K, dim]
    expansion = np.prod(centered**grid, axis=-1)              # [B, N, K]
    return np.transpose(expansion, [0, 2, 1])
  return basis


def make_basket_put_payoff(strikes, dtype=None):
  """`payoff_utils.py:27-97`: -> [num_samples, batch_size]."""
  strikes = np.asarray(strikes, dtype=dtype)

  def put_valuer(sample_paths, time_index):
    x = np.asarray(sample_paths, dtype=strikes.dtype)
    x = x[:, None] if x.ndim == 3 else np.transpose(x, [1, 0, 2, 3])
    sl = x[:, :, time_index, :]                               # [N, B, dim]
    return np.maximum(strikes - sl.mean(axis=-1), 0)
  return put_valuer


def _apply_discount(values, df, e):
  return (df[e + 1] / df[e]) * values


def least_square_mc(sample_paths, exercise_times, payoff_fn, basis_fn,
                    discount_factors=None, num_calibration_samples=None,
                    dtype=None, diagnostics=None):
  """`least_square_mc` (`lsm.py:128-295`) -> [batch_size] prices.

  `diagnostics` (oracle extension): a dict that receives, per exercise index e,
  the normal equations `lhs[e]` [B, K, K], `rhs[e]` [B, K] and `beta[e]`, and the
  final merged state `w` = cashflow + values [N, B]."""
  x = np.asarray(sample_paths, dtype=dtype)
  dtype = x.dtype
  exercise_times = np.asarray(exercise_times)
  T = exercise_times.shape[-1]
  if discount_factors is None:
    df = np.ones(exercise_times.shape, dtype=dtype)
  else:
    df = np.asarray(discount_factors, dtype=dtype)
  if df.ndim == 0:
    df = df.reshape(1, 1, 1)
  if df.ndim == 1:
    df = df.reshape(1, 1, -1)
  df = np.concatenate([np.ones(df.shape[:2] + (1,), dtype=dtype), df], axis=-1)
  df = np.transpose(df, [2, 0, 1])                            # [T+1, N|1, B|1]
  cashflow = payoff_fn(x, exercise_times[T - 1])              # [N, B]
  values = np.zeros_like(cashflow)
  calib = None if num_calibration_samples is None else np.arange(num_calibration_samples)
  e = T - 1
  while e > 0:
    t_idx = exercise_times[e - 1]
    ev = payoff_fn(x, t_idx)
    cont = _apply_discount(values + cashflow, df, e)
    design = basis_fn(x, t_idx)                               # [B, K, N]
    mask = ev > 0
    design_t = np.transpose(design, [0, 2, 1])                # [B, N, K]
    masked = np.where(mask.T[..., None], design_t, 0)
    if calib is None:
      sub, y = masked, cont
    else:
      sub, y = masked[:, calib], cont[calib]
    lhs = np.matmul(np.transpose(sub, [0, 2, 1]), sub)        # [B, K, K]
    K = lhs.shape[-1]
    pinv = np.stack([np.linalg.pinv(m, rcond=10 * K * np.finfo(dtype).eps) for m in lhs])
    rhs = np.matmul(np.transpose(sub, [0, 2, 1]), y.T[..., None])
    beta = np.matmul(pinv, rhs)
    if diagnostics is not None:
      diagnostics.setdefault('lhs', {})[e] = lhs
      diagnostics.setdefault('rhs', {})[e] = rhs[..., 0]
      diagnostics.setdefault('beta', {})[e] = beta[..., 0]
    expected = np.maximum(np.matmul(design_t, beta)[..., 0].T, 0)   # [N, B]
    upd = ev > expected
    new_values = np.where(upd, 0, cashflow + values)
    cashflow = np.where(upd, ev, 0).astype(dtype)
    values = _apply_discount(new_values, df, e).astype(dtype)
    e -= 1
  if diagnostics is not None:
    diagnostics['w'] = cashflow + values
  pv = _apply_discount(cashflow + values, df, 0)
  if num_calibration_samples is not None:
    pv = pv[num_calibration_samples:]
  return pv.mean(axis=0)
