"""Oracle (test infrastructure): pathwise tangents of the 1-d affine and of the Heston Euler
scheme.

The reference obtains sensitivities of sampled paths by differentiating the
Euler loop itself (`watch_params`, `models/euler_sampling.py:393-402, 467-510`
through `math/custom_loops.py:20-215`, used by the notebook
`examples/jupyter_notebooks/Monte_Carlo_Euler_Scheme.ipynb` cells 22-28 with
`tff.math.fwd_gradient`).  For dX = (a0 + a1 X) dt + (b0 + b1 X) dW one
`_euler_step` (`euler_sampling.py:513-537`) is
  X' = X + dt (a0 + a1 X) + (b0 + b1 X) dw,
so its forward-mode derivatives are
  Y' = Y + Y (dt a1 + b1 dw)                                   Y = dX/dX0
  V' = V + V (dt a1 + b1 dw) + dt (da0 + da1 X) + (db0 + db1 X) dw   V = dX/dtheta.
This module runs exactly that through the oracle's own sampler: the triple
(X, Y, V) is a 3-d Ito process driven by the FIRST normal of each step only, fed
with the draws of the 1-d process (columns 2 and 3 of `normal_draws` are zero),
so X is bit-identical to `oracle.euler.sample(1, ...)`.
"""
import numpy as np

from oracle import draws as draws_lib
from oracle import euler
from oracle import grid as grid_lib


def _p(param, t, dtype):
  if callable(param):
    return dtype.type(np.asarray(param(np.asarray([t], dtype=dtype)))[0])
  return dtype.type(param)


def sample_with_tangents(a0, a1, b0, b1, da0, da1, db0, db1, times, initial_state,
                         num_samples, random_type=None, seed=None, skip=0, time_step=None,
                         num_time_steps=None, times_grid=None, dtype=np.float64):
  """-> [num_samples, k, 3] with components [X, dX/dX0, dX/dtheta]."""
  dtype = np.dtype(dtype)
  times = np.asarray(times, dtype=dtype)
  all_times, _, _ = grid_lib.euler_grid(times, dtype=dtype, time_step=time_step,
                                        num_time_steps=num_time_steps, times_grid=times_grid,
                                        tolerance=None)
  steps = all_times.shape[0] - 1
  z = draws_lib.generate_mc_normal_draws(
      num_normal_draws=1, num_time_steps=steps, num_sample_paths=num_samples, batch_shape=(),
      random_type=draws_lib.RandomType.PSEUDO if random_type is None else random_type,
      dtype=dtype, seed=seed, skip=skip)                      # [steps, N, 1]
  draws = np.concatenate([z, np.zeros_like(z), np.zeros_like(z)], axis=-1)
  draws = np.transpose(draws, [1, 0, 2])                      # [N, steps, 3]

  def drift(t, s):
    x, y, v = s[..., 0], s[..., 1], s[..., 2]
    pa0, pa1 = _p(a0, t, dtype), _p(a1, t, dtype)
    return np.stack([pa0 + pa1 * x, y * pa1,
                     v * pa1 + (_p(da0, t, dtype) + _p(da1, t, dtype) * x)], axis=-1)

  def vol(t, s):
    x, y, v = s[..., 0], s[..., 1], s[..., 2]
    pb1 = _p(b1, t, dtype)
    col = np.stack([_p(b0, t, dtype) + pb1 * x, y * pb1,
                    v * pb1 + (_p(db0, t, dtype) + _p(db1, t, dtype) * x)], axis=-1)
    out = np.zeros(s.shape + (3,), dtype=dtype)
    out[..., 0] = col
    return out

  x0 = np.array([np.asarray(initial_state, dtype=dtype).reshape(-1)[0], 1.0, 0.0], dtype=dtype)
  return euler.sample(3, drift, vol, times, time_step=time_step, num_time_steps=num_time_steps,
                      initial_state=x0, times_grid=times_grid, normal_draws=draws, dtype=dtype)


def heston_with_tangents(mean_reversion, theta, volvol, rho, d_mean_reversion, d_theta, d_volvol,
                         d_rho, d_initial_state, times, initial_state, num_samples,
                         random_type=None, seed=None, skip=0, time_step=None, num_time_steps=None,
                         dtype=np.float64):
  """Heston Euler paths (closures of `heston/heston_model.py:143-173`) with the forward-mode
  derivatives of one `_euler_step` with respect to a scalar p: -> [num_samples, k, 4] with
  components [X, V, dX/dp, dV/dp].  With s = sqrt|V|, ds = sign(V) (dV/dp) / (2 s):
    dX/dp' = dX/dp - dt (dV/dp) / 2 + ds dw0
    dV/dp' = dV/dp + dt (dkappa (theta - V) + kappa (dtheta - dV/dp))
             + (dxi s + xi ds) (rho dw0 + rhobar dw1) + xi s (drho dw0 + drhobar dw1).
  The quadruple is a 4-d Ito process driven by the first TWO normals of each step, fed with
  the draws of the 2-d process: (X, V) is bit-identical to the plain Heston sampler."""
  dtype = np.dtype(dtype)
  times = np.asarray(times, dtype=dtype)
  all_times, _, _ = grid_lib.euler_grid(times, dtype=dtype, time_step=time_step,
                                        num_time_steps=num_time_steps, times_grid=None,
                                        tolerance=None)
  steps = all_times.shape[0] - 1
  z = draws_lib.generate_mc_normal_draws(
      num_normal_draws=2, num_time_steps=steps, num_sample_paths=num_samples, batch_shape=(),
      random_type=draws_lib.RandomType.PSEUDO if random_type is None else random_type,
      dtype=dtype, seed=seed, skip=skip)                      # [steps, N, 2]
  draws = np.transpose(np.concatenate([z, np.zeros_like(z)], axis=-1), [1, 0, 2])   # [N, steps, 4]
  k, th, xi, r = (lambda t, q=q: _p(q, t, dtype) for q in (mean_reversion, theta, volvol, rho))
  dk, dth, dxi, dr = (lambda t, q=q: _p(q, t, dtype) for q in (d_mean_reversion, d_theta, d_volvol, d_rho))

  def drift(t, s):
    v, vt = s[..., 1], s[..., 3]
    return np.stack([-v / 2, k(t) * (th(t) - v), -vt / 2,
                     dk(t) * (th(t) - v) + k(t) * (dth(t) - vt)], axis=-1)

  def vol(t, s):
    v, vt = s[..., 1], s[..., 3]
    sq = np.sqrt(np.abs(v))
    with np.errstate(divide='ignore', invalid='ignore'):
      ds = np.where(sq > 0, np.sign(v) * vt / (2 * sq), 0.0)
    rb = np.sqrt(1 - r(t)**2)
    drb = -(r(t) * dr(t)) / rb
    out = np.zeros(s.shape + (4,), dtype=dtype)
    out[..., 0, 0] = sq
    out[..., 1, 0] = xi(t) * r(t) * sq
    out[..., 1, 1] = xi(t) * rb * sq
    out[..., 2, 0] = ds
    amp = dxi(t) * sq + xi(t) * ds
    out[..., 3, 0] = amp * r(t) + xi(t) * sq * dr(t)
    out[..., 3, 1] = amp * rb + xi(t) * sq * drb
    return out

  x0 = np.concatenate([np.asarray(initial_state, dtype=dtype).reshape(-1)[:2],
                       np.asarray(d_initial_state, dtype=dtype).reshape(-1)[:2]])
  return euler.sample(4, drift, vol, times, time_step=time_step, num_time_steps=num_time_steps,
                      initial_state=x0, normal_draws=draws, dtype=dtype)
