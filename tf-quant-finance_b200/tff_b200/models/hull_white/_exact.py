"""Host-side tables of the one-factor Hull-White exact discretisation.

Restates, in numpy and for dim = 1, the closed-form integrals of
`models/hull_white/vector_hull_white.py:816-955` (`_exact_discretization_setup`,
`_compute_yt`, `_conditional_mean_x`, `_conditional_variance_x`, `_y_integral`,
`_ex_integral`, `_variance_int`).  They are evaluated once per grid point and
shipped to the device as the per-step coefficient table of TQF_MODEL_HW1F.
"""
import numpy as np
import torch


def forward_rate_fns(initial_discount_rate_fn, dtype):
  """(f, f') with f(0,t) = d/dt [r(t) t].

  The reference differentiates `-log P(0,t)` with forward-mode AD
  (`vector_hull_white.py:209-225, 298-300`).  Here: torch autograd when the
  callable accepts torch tensors, otherwise a numpy complex step (exact to
  rounding for analytic functions) with a central difference for f'.
  """
  dtype = np.dtype(dtype)

  def _torch_f(t, order):
    tt = torch.tensor(np.asarray(t, dtype=np.float64).reshape(-1),
                      dtype=torch.float64, requires_grad=True)
    r = initial_discount_rate_fn(tt)
    if not isinstance(r, torch.Tensor):
      raise TypeError('not a torch function')
    if r.dim() == tt.dim() + 1:
      r = r[..., 0]
    g = r * tt
    (f,) = torch.autograd.grad(g.sum(), tt, create_graph=order > 1,
                               allow_unused=True)
    if f is None:
      f = torch.zeros_like(tt)
    if order == 1:
      return f.detach().numpy()
    if not f.requires_grad:
      return np.zeros(tt.shape)
    (fp,) = torch.autograd.grad(f.sum(), tt, allow_unused=True)
    return np.zeros(tt.shape) if fp is None else fp.detach().numpy()

  def _np_f(t):
    z = np.asarray(t, dtype=np.float64).reshape(-1) + 1e-30j
    r = np.asarray(initial_discount_rate_fn(z))
    if r.ndim == z.ndim + 1:
      r = r[..., 0]
    return np.imag(r * z) / 1e-30

  use_torch = True
  try:
    _torch_f(np.array([0.5]), 1)
  except Exception:  # pylint: disable=broad-except
    use_torch = False

  def fwd(t):
    shape = np.shape(t)
    v = _torch_f(t, 1) if use_torch else _np_f(t)
    return np.asarray(v, dtype=dtype).reshape(shape)

  def fwd_grad(t):
    shape = np.shape(t)
    if use_torch:
      v = _torch_f(t, 2)
    else:
      tt = np.asarray(t, dtype=np.float64).reshape(-1)
      v = (_np_f(tt + 1e-5) - _np_f(tt - 1e-5)) / 2e-5
    return np.asarray(v, dtype=dtype).reshape(shape)
  return fwd, fwd_grad


def discount_rate(initial_discount_rate_fn, t, dtype):
  """r(t) of P(0,t) = exp(-r(t) t) as a numpy array shaped like `t`."""
  t = np.asarray(t, dtype=dtype)
  try:
    r = initial_discount_rate_fn(t)
    if isinstance(r, torch.Tensor):
      r = r.detach().cpu().numpy()
    r = np.asarray(r, dtype=dtype)
  except Exception:  # pylint: disable=broad-except
    r = initial_discount_rate_fn(torch.as_tensor(t))
    r = np.asarray(r.detach().cpu().numpy(), dtype=dtype)
  if r.ndim == t.ndim + 1:
    r = r[..., 0]
  return np.broadcast_to(r, t.shape)


class ExactTables:
  """Constant mean reversion `k`, piecewise-constant volatility `vol`."""

  def __init__(self, k, vol, dtype):
    self.dtype = np.dtype(dtype)
    self.k = self.dtype.type(k)
    self.vol = vol                               # PiecewiseConstantFunc, batch-free
    self.jumps = np.sort(np.asarray(vol.jump_locations(), dtype=self.dtype).reshape(-1))
    self.jump_vol = np.asarray(vol(self.jumps), dtype=self.dtype).reshape(-1)
    n = self.jumps.shape[0]
    self.padded_knots = np.concatenate(
        [np.zeros(1, dtype=self.dtype), self.jumps[:-1]])[:n]

  def _vol_at(self, t):
    return np.asarray(self.vol(t), dtype=self.dtype).reshape(np.shape(t))

  def _y_integral(self, t0, t, vol):
    k = self.k
    return (vol * vol) / (2 * k) * (np.exp(2 * k * t) - np.exp(2 * k * t0))

  def _ex_integral(self, t0, t, vol, y_t0):
    k = self.k
    value = (np.exp(k * t) - np.exp(k * t0) + np.exp(2 * k * t0) *
             (np.exp(-k * t) - np.exp(-k * t0)))
    return value * vol**2 / (2 * k * k) + y_t0 * (np.exp(-k * t0) - np.exp(-k * t)) / k

  def _cum_at_knots(self):
    between = self._y_integral(self.padded_knots, self.jumps, self.jump_vol)
    return np.concatenate([np.zeros(1, dtype=self.dtype), np.cumsum(between)])

  def y_t(self, t):
    t = np.asarray(t, dtype=self.dtype)
    idx = np.searchsorted(self.jumps, t, side='left')
    vn = np.concatenate([np.zeros(1, dtype=self.dtype), self.jumps])
    y = self._y_integral(vn[idx], t, self._vol_at(t)) + self._cum_at_knots()[idx]
    return np.exp(-2 * self.k * t) * y

  def conditional_mean_x(self, t):
    t = np.asarray(t, dtype=self.dtype)
    idx = np.searchsorted(self.jumps, t, side='left')
    vn = np.concatenate([np.zeros(1, dtype=self.dtype), self.jumps])
    y_at = self._cum_at_knots()
    ex_between = self._ex_integral(self.padded_knots, self.jumps, self.jump_vol,
                                   y_at[:-1])
    ex_at = np.concatenate([np.zeros(1, dtype=self.dtype), np.cumsum(ex_between)])
    ex = self._ex_integral(vn[idx], t, self._vol_at(t), y_at[idx]) + ex_at[idx]
    return (ex[1:] - ex[:-1]) * np.exp(-self.k * t[1:])

  def conditional_variance_x(self, t):
    t = np.asarray(t, dtype=self.dtype)
    idx = np.searchsorted(self.jumps, t, side='left')
    vn = np.concatenate([np.zeros(1, dtype=self.dtype), self.jumps])
    var = self._y_integral(vn[idx], t, self._vol_at(t)) + self._cum_at_knots()[idx]
    return (var[1:] - var[:-1]) * np.exp(-2 * self.k * t[1:])
