"""Gaussian HJM (`models/hjm/gaussian_hjm.py:32-465`) on the B200 path engine:
deterministic volatility (constant per factor or a `PiecewiseConstantFunc`), state
`x` only (F draws per step), `y(t)` in closed form (`state_y`, 316-373), discount
factors by the left-point rule (451-456)."""
import numpy as np
import torch

from tff_b200 import _tensor
from tff_b200.math import piecewise
from tff_b200.models.hjm import quasi_gaussian_hjm


class GaussianHJM(quasi_gaussian_hjm.QuasiGaussianHJM):
  """`GaussianHJM(dim, mean_reversion, volatility, initial_discount_rate_fn,
  corr_matrix=None, dtype=None, name=None)`, `dim` <= 3 factors."""

  _DRAWS_PER_STEP_IS_STATE_DIM = False
  _RIGHT_POINT_DISCOUNTING = False

  def __init__(self, dim, mean_reversion, volatility, initial_discount_rate_fn,
               corr_matrix=None, dtype=None, name=None):
    dt_ = _tensor.np_dtype(dtype, np.float32)
    f = int(dim)
    if isinstance(volatility, piecewise.PiecewiseConstantFunc):
      jumps = np.asarray(volatility.jump_locations(), dtype=dt_)
      values = np.asarray(volatility.values(), dtype=dt_)
      self._vol_jumps = np.broadcast_to(jumps.reshape(-1, jumps.shape[-1]), (f, jumps.shape[-1]))
      self._vol_values = np.broadcast_to(values.reshape(-1, values.shape[-1]),
                                         (f, values.shape[-1]))
    elif callable(volatility):
      raise NotImplementedError(
          'GaussianHJM: pass the volatility as a tensor or a PiecewiseConstantFunc (the closed '
          'form of y(t) needs its jump locations, gaussian_hjm.py:455-465).')
    else:
      self._vol_jumps = np.zeros((f, 0), dt_)
      self._vol_values = _tensor.to_numpy(volatility, dt_).reshape(f, 1)
    super().__init__(dim, mean_reversion, lambda t, r: self._sigma_at(t),
                     initial_discount_rate_fn, corr_matrix=corr_matrix, validate_args=True,
                     dtype=dt_, name=name or 'gaussian_hjm_model')
    self._dim = f                                       # gaussian_hjm.py:161-162

  def _sigma_at(self, t):
    t = float(t)
    return np.array([self._vol_values[i][np.searchsorted(self._vol_jumps[i], t, side='left')]
                     for i in range(self._factors)], dtype=self._dtype)

  def _sigma(self, t):
    return self._sigma_at(t)

  def state_y(self, t, name=None):
    """y_ij(t) = e^{-(k_i + k_j) t} int_0^t rho_ij sigma_i(u) sigma_j(u) e^{(k_i + k_j) u} du
    as a numpy array `[F, F, len(t)]` (`gaussian_hjm.py:316-373`)."""
    del name
    t = np.asarray(_tensor.to_numpy(t, self._dtype), dtype=np.float64).reshape(-1)
    f = self._factors
    k = self._mean_reversion.astype(np.float64)
    out = np.zeros((f, f, t.shape[0]))
    for i in range(f):
      for j in range(f):
        c = k[i] + k[j]
        knots = np.union1d(self._vol_jumps[i], self._vol_jumps[j]).astype(np.float64)
        edges = np.concatenate([[0.0], knots, [np.inf]])
        mids = np.where(np.isinf(edges[1:]), edges[:-1] + 1.0, 0.5 * (edges[:-1] + edges[1:]))
        si = self._vol_values[i][np.searchsorted(self._vol_jumps[i], mids, side='left')]
        sj = self._vol_values[j][np.searchsorted(self._vol_jumps[j], mids, side='left')]
        # piece p covers (edges[p], edges[p + 1]]: its part below t, scaled by e^{-c t}
        lo = np.minimum(edges[None, :-1], t[:, None])
        hi = np.minimum(edges[None, 1:], t[:, None])
        piece = (np.exp(c * (hi - t[:, None])) - np.exp(c * (lo - t[:, None]))) / c
        out[i, j] = float(self._rho[i, j]) * (piece * (si * sj)[None, :]).sum(-1)
    return out.astype(self._dtype)

  def _drift_a0(self, all_times, y_entries):
    del y_entries
    return np.transpose(self.state_y(all_times[1:]).sum(1))          # [S, F]: sum_j y_ij(t_{s+1})

  def _y_at(self, times, y_simulated):
    del y_simulated
    return np.transpose(self.state_y(times), (2, 0, 1))               # [k, F, F]

  def discount_bond_price(self, state, times, maturities, name=None):
    """P(t, T) given x(t) (`gaussian_hjm.py:375-411`): `state` [n, F], `times` [n],
    `maturities` [n] -> numpy [n] (a closed form evaluated on the host)."""
    del name
    dt_ = self._dtype
    x = _tensor.to_numpy(state, dt_).reshape(-1, self._factors)
    times = _tensor.to_numpy(times, dt_).reshape(-1)
    maturities = _tensor.to_numpy(maturities, dt_).reshape(-1)
    y = np.transpose(self.state_y(times), (2, 0, 1))
    a, g = self._bond_tables(times, maturities[None, :], y)          # [1, n], [1, n, F]
    return (a[0] * np.exp(-(g[0] * x).sum(-1))).astype(dt_)
