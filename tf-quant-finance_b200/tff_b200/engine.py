"""Host driver of the fused B200 path engine (libtqf `tqf_plan_*`).

A *model spec* turns model parameters into the per-step coefficient table the
kernels consume (parameters are evaluated at `times[i + 1]`, the END of each
step, exactly as `_euler_step` does -- `models/euler_sampling.py:520`).  A
`Plan` owns the device-resident tables; `Plan.paths` materialises states,
`Plan.price` reduces payoffs in-kernel.
"""
import ctypes as C
import enum

import numpy as np
import torch

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200.math import random
from tff_b200.math.random import philox
from tff_b200.math.random import sobol


# ----------------------------------------------------------------- specs ----
def _eval_param(p, t, dtype):
  """Scalar model parameter at the array of times `t` (host)."""
  if callable(p):
    return np.asarray(_tensor.to_numpy(p(t), dtype), dtype=dtype).reshape(t.shape)
  return np.broadcast_to(np.asarray(_tensor.to_numpy(p, dtype)).reshape(()),
                         t.shape).astype(dtype)


class ModelSpec:
  """Base class: a model the device kernels can step."""
  kind = None
  dim = None
  num_factors = None
  num_coef = None

  def coef_table(self, all_times, dtype):
    """float64 [S, num_coef]; values exactly representable in `dtype`."""
    raise NotImplementedError

  def constant_key(self):
    """A hashable description of the model when ALL its parameters are plain numbers
    (nothing a later call could evaluate differently), else None.  Lets a pricing
    entry point recognise a repeated call without rebuilding the tables."""
    return None

  @staticmethod
  def _dt_columns(all_times, dtype):
    t = np.asarray(all_times, dtype=dtype)
    dt = (t[1:] - t[:-1]).astype(dtype)
    return t[1:], dt, np.sqrt(dt).astype(dtype)


class AffineSpec1F(ModelSpec):
  """dX = (a0(t) + a1(t) X) dt + (b0(t) + b1(t) X) dW.

  When both state coefficients are the constant 0 (additive noise: the log-space
  GBM of the Monte-Carlo notebook, configs C1 / C5) the Euler step
  `(x + dt a0) + b0 (z sqrt_dt)` does not read the state in its increments; it then
  runs as `x' = (x + B) + C z` with the per-step constants `B = dt a0`,
  `C = b0 sqrt_dt` formed once on the host (TQF_MODEL_LINEAR_1F: 2 fused
  multiply-adds per path-step instead of 7 FP64 instructions; the result differs
  from the reference grouping by one rounding of `b0 sqrt_dt z`)."""
  dim, num_factors = 1, 1

  def __init__(self, a0, a1, b, b1=0.0):
    self.a0, self.a1, self.b, self.b1 = a0, a1, b, b1
    self.additive = all(
        not callable(p) and np.ndim(p) == 0 and float(p) == 0.0 for p in (a1, b1))
    self.kind = _lib.MODEL_LINEAR_1F if self.additive else _lib.MODEL_AFFINE_1F
    self.num_coef = 5 if self.additive else 6

  def constant_key(self):
    ps = (self.a0, self.a1, self.b, self.b1)
    if any(callable(p) or np.ndim(p) != 0 for p in ps):
      return None
    return ('affine1f',) + tuple(float(p) for p in ps)

  def general_table(self, all_times, dtype):
    """The six columns dt, sqrt_dt, a0, a1, b0, b1 of TQF_MODEL_AFFINE_1F."""
    t, dt, sq = self._dt_columns(all_times, dtype)
    cols = [dt, sq, _eval_param(self.a0, t, dtype), _eval_param(self.a1, t, dtype),
            _eval_param(self.b, t, dtype), _eval_param(self.b1, t, dtype)]
    return np.stack(cols, -1).astype(np.float64)

  def coef_table(self, all_times, dtype):
    if not self.additive:
      return self.general_table(all_times, dtype)
    t, dt, sq = self._dt_columns(all_times, dtype)
    a0 = _eval_param(self.a0, t, dtype)
    b0 = _eval_param(self.b, t, dtype)
    cols = [dt, sq, np.ones_like(dt), (dt * a0).astype(dtype), (b0 * sq).astype(dtype)]
    return np.stack(cols, -1).astype(np.float64)


class TangentAffineSpec1F(ModelSpec):
  """`AffineSpec1F` carrying the pathwise tangents `dX/dX0` and `dX/dtheta`
  alongside the path: the forward-mode Jacobians the reference obtains with
  `watch_params` (`euler_sampling.py:393-402, 467-510`,
  `math/custom_loops.py:20-215`).  `theta` is one scalar parameter; `da0`,
  `da1`, `db`, `db1` are the derivatives of the four coefficient functions with
  respect to it (scalars or callables of an array of times).  The device state
  is `[X, dX/dX0, dX/dtheta]`; the user-facing dimension stays 1 and the draws
  are those of the plain 1-d process."""
  kind, dim, num_factors, num_coef = _lib.MODEL_AFFINE_1F_TANGENT, 3, 1, 10
  user_dim = 1
  X, D_INITIAL, D_THETA = 0, 1, 2      # state components

  def __init__(self, a0, a1, b, b1=0.0, da0=0.0, da1=0.0, db=0.0, db1=0.0):
    self.p = (a0, a1, b, b1, da0, da1, db, db1)

  def coef_table(self, all_times, dtype):
    t, dt, sq = self._dt_columns(all_times, dtype)
    cols = [dt, sq] + [_eval_param(q, t, dtype) for q in self.p]
    return np.stack(cols, -1).astype(np.float64)

  @staticmethod
  def extend_initial_state(x0):
    return np.concatenate([np.asarray(x0).reshape(-1)[:1], [1.0, 0.0]])


class MilsteinSpec1F(ModelSpec):
  """The Milstein scheme (`models/milstein_sampling.py:565-575`) of a 1-d process
  whose Euler spec is affine (`AffineSpec1F`, `GbmSpec1F` or a probed pair of
  plain callables): the volatility gradient of `b0(t) + b1(t) x` is `b1(t)`."""
  kind, dim, num_factors, num_coef = _lib.MODEL_MILSTEIN_1F, 1, 1, 6

  def __init__(self, euler_spec):
    if euler_spec.dim != 1 or not (isinstance(euler_spec, AffineSpec1F) or euler_spec.kind in (
        _lib.MODEL_AFFINE_1F, _lib.MODEL_GBM_1F)):
      raise NotImplementedError(
          'The B200 Milstein kernel covers 1-d processes with affine drift and '
          'volatility (affine_closures, gbm_closures, GeometricBrownianMotion, or '
          'plain callables that are affine in the state). There is no CPU fallback.')
    self.euler_spec = euler_spec

  def coef_table(self, all_times, dtype):
    if isinstance(self.euler_spec, AffineSpec1F):
      tab = self.euler_spec.general_table(all_times, dtype)
    else:
      tab = self.euler_spec.coef_table(all_times, dtype)
    if self.euler_spec.kind == _lib.MODEL_GBM_1F:      # dt, sqrt_dt, mu, sigma
      zero = np.zeros_like(tab[:, 0])
      tab = np.stack([tab[:, 0], tab[:, 1], zero, tab[:, 2], zero, tab[:, 3]], -1)
    return np.ascontiguousarray(tab, dtype=np.float64)


class ProbedAffineSpec(ModelSpec):
  """An arbitrary Python (drift_fn, volatility_fn) pair that turns out to be
  affine: a(t, x) = a0(t) + A1(t) x and S(t, x) = B0(t) (+ B1(t) x for dim 1).

  The callables are evaluated ON THE HOST at every grid time for a few probe
  states; the coefficients are solved from the origin and the unit vectors and
  verified on check points of both signs, of magnitudes 1e-3 .. 1e3 and around
  the initial state.  A pair that is not affine there, or that is not finite on
  a probe (1 / x, log x), is rejected -- the CUDA
  kernels cannot run Python, and there is no CPU fallback."""

  def __init__(self, dim, drift_fn, volatility_fn):
    if dim < 1 or dim > 4:
      raise NotImplementedError(
          'generic drift/volatility callables are supported for dim <= 4')
    self.dim = self.num_factors = int(dim)
    self.kind = _lib.MODEL_AFFINE_1F if dim == 1 else _lib.MODEL_AFFINE_ND
    self.num_coef = 6 if dim == 1 else 2 + dim + 2 * dim * dim
    self.drift_fn, self.volatility_fn = drift_fn, volatility_fn

  def _call(self, fn, t, x, dtype, want_matrix):
    """fn(t, x) with torch tensors first, numpy as a second attempt."""
    d = self.dim
    errors = []
    for mode in ('torch', 'numpy'):
      try:
        if mode == 'torch':
          tt = torch.tensor(float(t), dtype=_tensor.torch_dtype(dtype))
          xx = torch.as_tensor(np.ascontiguousarray(x, dtype=dtype))
          out = fn(tt, xx)
          if isinstance(out, torch.Tensor):
            out = out.detach().cpu().numpy()
        else:
          out = fn(dtype.type(t), np.ascontiguousarray(x, dtype=dtype))
        out = np.asarray(out, dtype=np.float64)
        shape = (x.shape[0], d, d) if want_matrix else (x.shape[0], d)
        return np.broadcast_to(out, shape)
      except Exception as e:  # pylint: disable=broad-except
        errors.append('%s: %r' % (mode, e))
    raise NotImplementedError(
        'could not evaluate the drift/volatility callable on the host '
        '(tried torch and numpy inputs): ' + '; '.join(errors))

  def coef_table(self, all_times, dtype):
    dtype = np.dtype(dtype)
    d = self.dim
    t, dt, sq = self._dt_columns(all_times, dtype)
    rs = np.random.RandomState(12345)
    # probes: origin and unit vectors (the coefficients are solved from these), then
    # check points on which the affine form must reproduce the callable: both signs,
    # small and large magnitudes, and a cloud around the initial state -- a callable
    # that is only affine on part of the state space (relu, abs, clamp, a local-vol
    # cut-off) must not pass
    checks = [rs.uniform(0.5, 2.0, (2, d)), rs.uniform(-2.0, -0.5, (2, d)),
              rs.uniform(-1.0, 1.0, (2, d)) * 1e-3, rs.uniform(-1.0, 1.0, (2, d)) * 1e3]
    x0 = getattr(self, 'initial_state_hint', None)
    if x0 is not None:
      x0 = np.asarray(x0, dtype=np.float64).reshape(-1)[:d]
      if x0.shape[0] == d and np.all(np.isfinite(x0)):
        spread = np.maximum(np.abs(x0), 1.0)
        checks.append(x0 + spread * rs.uniform(-1.0, 1.0, (4, d)))
    probes = np.concatenate([np.zeros((1, d)), np.eye(d)] + checks)
    rows = []
    for i in range(t.shape[0]):
      a = self._call(self.drift_fn, t[i], probes, dtype, False)          # [P, d]
      s_ = self._call(self.volatility_fn, t[i], probes, dtype, True)     # [P, d, d]
      if not (np.all(np.isfinite(a)) and np.all(np.isfinite(s_))):
        raise NotImplementedError(
            'The drift/volatility callable returns non-finite values on the probe states '
            '(origin, unit vectors, points of both signs) at t={}: it is not an affine '
            'function of the state on the whole state space, which is what the B200 path '
            'engine runs for plain Python callables. There is no CPU fallback.'.format(
                float(t[i])))
      a0 = a[0]
      a1 = (a[1:1 + d] - a0).T                                           # A1[i][j]
      b0 = s_[0]
      b1 = s_[1:1 + d] - b0                                              # [j][., .]
      for c in range(d + 1, probes.shape[0]):
        x = probes[c]
        pred_a = a0 + a1 @ x
        pred_s = b0 + np.tensordot(x, b1, axes=(0, 0))
        scale = 1.0 + np.abs(a[c]).max() + np.abs(s_[c]).max()
        if (np.abs(pred_a - a[c]).max() > 1e-6 * scale or
            np.abs(pred_s - s_[c]).max() > 1e-6 * scale):
          raise NotImplementedError(
              'The B200 path engine runs affine drift/volatility callables '
              '(a = a0(t) + A1(t) x, S = B(t)) and the closures of its model '
              'classes; this callable pair is not affine in the state at t={}. '
              'There is no CPU fallback.'.format(float(t[i])))
      if d == 1:
        rows.append([dt[i], sq[i], a0[0], a1[0, 0], b0[0, 0], b1[0, 0, 0]])
      else:
        if np.abs(b1).max() > 1e-9 * (1.0 + np.abs(b0).max()):
          raise NotImplementedError(
              'state-dependent volatility callables are supported for dim 1 '
              'only (dim > 1: use the model classes); no CPU fallback.')
        rows.append(np.concatenate([[dt[i], sq[i]], a0, a1.reshape(-1), b0.reshape(-1)]))
    table = np.asarray(rows, dtype=np.float64).reshape(t.shape[0], self.num_coef)
    return table.astype(dtype).astype(np.float64)


class GbmSpec1F(ModelSpec):
  """dX = mu(t) X dt + sigma(t) X dW."""
  kind, dim, num_factors, num_coef = _lib.MODEL_GBM_1F, 1, 1, 4

  def __init__(self, mean, volatility):
    self.mean, self.volatility = mean, volatility

  def for_batch(self, index, batch_shape):
    """The spec of one element of a batch of GBMs (parameters of shape
    `batch_shape + [1]`, broadcastable)."""
    def pick(p):
      if callable(p) or np.ndim(p) == 0:
        return p
      arr = np.broadcast_to(np.asarray(p), tuple(batch_shape) + (1,))
      return arr[tuple(index)][0]
    return GbmSpec1F(pick(self.mean), pick(self.volatility))

  def coef_table(self, all_times, dtype):
    t, dt, sq = self._dt_columns(all_times, dtype)
    cols = [dt, sq, _eval_param(self.mean, t, dtype),
            _eval_param(self.volatility, t, dtype)]
    return np.stack(cols, -1).astype(np.float64)


class LinearSpec1F(ModelSpec):
  """x' = A_i x + B_i + C_i z with explicit per-step tables (exact OU step)."""
  kind, dim, num_factors, num_coef = _lib.MODEL_LINEAR_1F, 1, 1, 5

  def __init__(self, table_fn):
    self._table_fn = table_fn          # all_times, dtype -> (A, B, C) arrays

  def coef_table(self, all_times, dtype):
    _, dt, sq = self._dt_columns(all_times, dtype)
    a, b, c = self._table_fn(np.asarray(all_times, dtype=dtype), dtype)
    cols = [dt, sq, np.asarray(a, dtype), np.asarray(b, dtype), np.asarray(c, dtype)]
    return np.stack(cols, -1).astype(np.float64)


class HestonEulerSpec(ModelSpec):
  """Heston closures (`heston/heston_model.py:143-173`), state [log S, V].

  Columns: sqrt_dt, -dt/2, dt kappa, theta, volvol rho sqrt_dt,
  volvol sqrt(1 - rho^2) sqrt_dt -- the per-step products the kernel needs,
  formed once here in `dtype` (parameters at t_{i+1})."""
  kind, dim, num_factors, num_coef = _lib.MODEL_HESTON_EULER, 2, 2, 6

  def __init__(self, mean_reversion, theta, volvol, rho):
    self.mean_reversion, self.theta = mean_reversion, theta
    self.volvol, self.rho = volvol, rho

  def coef_table(self, all_times, dtype):
    t, dt, sq = self._dt_columns(all_times, dtype)
    ty = np.dtype(dtype).type
    kappa = _eval_param(self.mean_reversion, t, dtype)
    theta = _eval_param(self.theta, t, dtype)
    volvol = _eval_param(self.volvol, t, dtype)
    rho = _eval_param(self.rho, t, dtype)
    c4 = ((volvol * rho).astype(dtype) * sq).astype(dtype)
    c5 = ((volvol * np.sqrt(ty(1) - rho**2).astype(dtype)).astype(dtype) * sq).astype(dtype)
    cols = [sq, (ty(-0.5) * dt).astype(dtype), (dt * kappa).astype(dtype), theta, c4, c5]
    return np.stack(cols, -1).astype(np.float64)


class TangentHestonSpec(ModelSpec):
  """`HestonEulerSpec` carrying the pathwise tangents of `(X, V)` with respect to ONE
  scalar `p` alongside the path (the forward-mode sensitivities the reference obtains by
  differentiating the Euler loop with `watch_params`, `euler_sampling.py:393-402`).
  `d_mean_reversion`, `d_theta`, `d_volvol`, `d_rho` are the derivatives of the four
  parameters with respect to `p`, `d_initial_state` those of `[X_0, V_0]` (so vega-type
  sensitivities to `V_0` and delta come from the same kernel).  Device state
  `[X, V, dX/dp, dV/dp]`; the user-facing dimension stays 2 and the draws are those of
  the plain Heston process."""
  kind, dim, num_factors, num_coef = _lib.MODEL_HESTON_TANGENT, 4, 2, 12
  user_dim = 2
  X, V, D_X, D_V = 0, 1, 2, 3          # state components

  def __init__(self, mean_reversion, theta, volvol, rho, d_mean_reversion=0.0, d_theta=0.0,
               d_volvol=0.0, d_rho=0.0, d_initial_state=(0.0, 0.0)):
    self.p = (mean_reversion, theta, volvol, rho)
    self.d = (d_mean_reversion, d_theta, d_volvol, d_rho)
    self.d_initial_state = tuple(float(v) for v in d_initial_state)

  def coef_table(self, all_times, dtype):
    t, dt, sq = self._dt_columns(all_times, dtype)
    ty = np.dtype(dtype).type
    kappa, theta, volvol, rho = (_eval_param(q, t, dtype) for q in self.p)
    dk, dth, dxi, drho = (_eval_param(q, t, dtype) for q in self.d)
    rhobar = np.sqrt(ty(1) - rho**2).astype(dtype)
    drhobar = (-(rho * drho) / rhobar).astype(dtype)
    cols = [dt, sq, kappa, theta, volvol, rho, rhobar, dk, dth, dxi, drho, drhobar]
    return np.stack(cols, -1).astype(np.float64)

  def extend_initial_state(self, x0):
    x0 = np.asarray(x0).reshape(-1)[:2]
    return np.concatenate([x0, np.asarray(self.d_initial_state, dtype=x0.dtype)])


class MvGbmSpec(ModelSpec):
  """Correlated multi-asset GBM closures
  (`geometric_brownian_motion/multivariate_geometric_brownian_motion.py:130-151`):
  a_i = mu_i x_i, S_ij = sigma_i x_i L_ij with L = cholesky(corr).  Table
  columns: dt, sqrt_dt; the Cholesky factor and (mu, sigma) travel as kernel
  parameters."""
  kind, num_coef = _lib.MODEL_MVGBM, 2

  def __init__(self, means, volatilities, corr_matrix, dim, exact_log=False):
    self.dim = self.num_factors = int(dim)
    self.means, self.volatilities, self.corr_matrix = means, volatilities, corr_matrix
    # exact_log: the state is log x and the step adds the exact log-normal
    # increment (means - vols^2/2) dt + sqrt(dt) vols (L z)
    self.exact_log = bool(exact_log)

  def coef_table(self, all_times, dtype):
    _, dt, sq = self._dt_columns(all_times, dtype)
    return np.stack([dt, sq], -1).astype(np.float64)

  def device_arrays(self, dtype):
    d = self.dim
    mu = np.broadcast_to(np.asarray(self.means, dtype=dtype), (d,))
    sg = np.broadcast_to(np.asarray(self.volatilities, dtype=dtype), (d,))
    if self.exact_log:
      mu = (mu - sg**2 / 2).astype(dtype)
    if self.corr_matrix is None:
      chol = np.eye(d, dtype=dtype)
    else:
      chol = np.linalg.cholesky(np.asarray(self.corr_matrix, dtype=dtype)).astype(dtype)
    mat = np.ascontiguousarray(chol, dtype=np.float64)
    vec = np.ascontiguousarray(np.stack([mu, sg]), dtype=np.float64)
    return mat, vec


# --------------------------------------------------------------- payoffs ----
class Payoff:
  """A payoff reduced in-kernel (the `tf.reduce_mean(tf.nn.relu(...))` tails
  of the reference's callers, e.g. `hull_white/swaption.py:310-311`)."""

  def __init__(self, kind, strike=0.0, barrier=0.0, component=0, log_state=False,
               scale=1.0, tangent=0, brownian_bridge=False):
    """`brownian_bridge` (barrier payoffs): continuous monitoring -- the payoff is
    weighted by the probability that the Brownian bridge between consecutive grid
    points did not touch the barrier (`tff.black_scholes.brownian_bridge_single`,
    `black_scholes/brownian_bridge.py:118-196`), accumulated step by step in the
    fused kernel."""
    self.kind, self.strike, self.barrier = kind, float(strike), float(barrier)
    self.component, self.log_state, self.scale = int(component), bool(log_state), float(scale)
    self.tangent = int(tangent)
    self.brownian_bridge = bool(brownian_bridge)

  def key(self):
    return (self.kind, self.strike, self.barrier, self.component, self.log_state, self.scale,
            self.tangent, self.brownian_bridge)

  def desc(self):
    d = _lib.PayoffDesc()
    d.kind, d.component = self.kind, self.component
    d.tangent_component = self.tangent
    d.brownian_bridge = int(self.brownian_bridge)
    d.transform = _lib.TRANSFORM_EXP if self.log_state else _lib.TRANSFORM_NONE
    d.strike, d.barrier, d.scale = self.strike, self.barrier, self.scale
    return d


def european_call(strike, **kw):
  return Payoff(_lib.PAYOFF_CALL, strike=strike, **kw)


def european_put(strike, **kw):
  return Payoff(_lib.PAYOFF_PUT, strike=strike, **kw)


def up_and_out_call(strike, barrier, **kw):
  return Payoff(_lib.PAYOFF_UP_OUT_CALL, strike=strike, barrier=barrier, **kw)


def up_and_out_put(strike, barrier, **kw):
  return Payoff(_lib.PAYOFF_UP_OUT_PUT, strike=strike, barrier=barrier, **kw)


def down_and_out_put(strike, barrier, **kw):
  return Payoff(_lib.PAYOFF_DOWN_OUT_PUT, strike=strike, barrier=barrier, **kw)


def down_and_out_call(strike, barrier, **kw):
  return Payoff(_lib.PAYOFF_DOWN_OUT_CALL, strike=strike, barrier=barrier, **kw)


def identity(**kw):
  return Payoff(_lib.PAYOFF_IDENTITY, **kw)


def european_call_tangent(strike, tangent, **kw):
  """Pathwise derivative of `european_call`: `1{f > K} f'(X_T) T_T`, with `T`
  the state component `tangent` of a tangent-carrying model
  (`TangentAffineSpec1F.D_INITIAL` -> delta-like, `.D_THETA` -> vega-like)."""
  return Payoff(_lib.PAYOFF_CALL_TANGENT, strike=strike, tangent=tangent, **kw)


def european_put_tangent(strike, tangent, **kw):
  """Pathwise derivative of `european_put`: `-1{K > f} f'(X_T) T_T`."""
  return Payoff(_lib.PAYOFF_PUT_TANGENT, strike=strike, tangent=tangent, **kw)


# ----------------------------------------------------------- record plan ----
def record_plan(keep_mask, num_requested_times):
  """Replays the slot bookkeeping of `_while_loop`
  (`models/euler_sampling.py:405-464`, `_euler_step` 531-535).

  Returns (num_steps_to_execute, record_slot int32 [steps + 1]) where entry 0
  refers to the initial state and entry s + 1 to the state after step s.
  """
  keep_mask = np.asarray(keep_mask, dtype=bool)
  steps_num = keep_mask.shape[0] - 1
  k = int(num_requested_times)
  last_writer = {0: 0}                    # slot -> entry (0 = initial state)
  written = int(keep_mask[0])
  i = 0
  while i < steps_num and written < k:
    last_writer[written] = i + 1
    written += int(keep_mask[i + 1])
    i += 1
  slots = np.full(i + 1, -1, dtype=np.int32)
  for slot, entry in last_writer.items():
    if slot < k:
      slots[entry] = slot
  return i, slots


# ------------------------------------------------------------------ plan ----
class RngSpec:
  """Resolved random-number configuration of one sampling call."""

  def __init__(self, random_type=None, seed=None, skip=0, normal_draws=None,
               unit_stride=1, unit_offset=0, total_units=None):
    # batched calls: path p draws unit p * unit_stride + unit_offset
    self.unit_stride, self.unit_offset = int(unit_stride), int(unit_offset)
    self.total_units = total_units
    rt = random.RandomType.PSEUDO if random_type is None else random_type
    if isinstance(rt, enum.Enum):
      rt = random.RandomType(rt.value)
    self.random_type = rt
    self.seed, self.skip = seed, int(skip or 0)
    self.normal_draws = normal_draws
    self.antithetic = rt in (random.RandomType.PSEUDO_ANTITHETIC,
                             random.RandomType.STATELESS_ANTITHETIC)
    if normal_draws is not None:
      self.type = _lib.RNG_DRAWS
      self.antithetic = False
    elif rt in (random.RandomType.PSEUDO, random.RandomType.PSEUDO_ANTITHETIC):
      self.type = _lib.RNG_PHILOX
      self.key, self.counter = philox.stateful_key_counter(seed)
    elif rt in (random.RandomType.STATELESS,
                random.RandomType.STATELESS_ANTITHETIC):
      if seed is None:
        raise ValueError('`seed` should be specified if the `random_type` is '
                         '`STATELESS` or `STATELESS_ANTITHETIC`')
      self.type = _lib.RNG_PHILOX
      self.key, self.counter = philox.stateless_key_counter(seed)
    elif rt == random.RandomType.SOBOL:
      self.type = _lib.RNG_SOBOL
    elif rt in (random.RandomType.HALTON, random.RandomType.HALTON_RANDOMIZED):
      raise NotImplementedError(
          'HALTON sequences are outside the B200 hot path; supported: PSEUDO, '
          'STATELESS, PSEUDO_ANTITHETIC, STATELESS_ANTITHETIC, SOBOL.')
    else:
      raise NotImplementedError(
          'Only STATELESS, PSEUDO, PSEUDO_ANTITHETIC, STATELESS_ANTITHETIC,  '
          'HALTON, HALTON_RANDOMIZED, and SOBOL random types are currently '
          'supported. Supplied: {}'.format(random_type))


class Plan:
  """Device-resident tables of one (model, grid, rng) configuration."""

  def __init__(self, spec, all_times, num_steps, x0, rng, num_samples, dtype, x0_paths=None,
               table=None):
    """`x0_paths`: optional per-path initial states `[num_samples, dim]`
    (anything convertible to a tensor); `x0` is then only a placeholder.
    `table`: the precomputed `spec.coef_table(all_times, dtype)[:num_steps]`."""
    self.spec, self.rng = spec, rng
    self.cached = False
    self.dtype = np.dtype(dtype)
    self.num_samples = int(num_samples)
    self.num_steps = int(num_steps)
    all_times = np.asarray(all_times, dtype=self.dtype)
    self.num_steps_total = all_times.shape[0] - 1
    if rng.antithetic and self.num_samples % 2 != 0:
      raise ValueError('First dimension of `sample_shape` should be even for '
                       'PSEUDO_ANTITHETIC random type')
    self.units = self.num_samples // 2 if rng.antithetic else self.num_samples

    if table is None:
      table = spec.coef_table(all_times, self.dtype)[:self.num_steps]
    table = np.ascontiguousarray(table, dtype=np.float64)
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=self.dtype).reshape(-1),
                              dtype=np.float64)
    if x0.shape[0] != spec.dim:
      raise ValueError('initial state must have {} components'.format(spec.dim))
    self._keep = [table, x0]

    m = _lib.ModelDesc()
    m.kind, m.dtype = spec.kind, _tensor.tqf_dtype(self.dtype)
    m.dim, m.num_factors = spec.dim, spec.num_factors
    m.num_steps, m.num_steps_total = self.num_steps, self.num_steps_total
    m.num_coef = spec.num_coef
    m.reserved = int(getattr(spec, 'exact_log', False))
    m.coef = table.ctypes.data
    m.x0 = x0.ctypes.data
    if x0_paths is not None:
      xp = torch.as_tensor(np.array(_tensor.to_numpy(x0_paths, self.dtype), order='C', copy=True),
                           device=_tensor.device()).contiguous()
      if tuple(xp.shape) != (self.num_samples, spec.dim):
        raise ValueError('per-path initial states must have shape {} but have {}'.format(
            (self.num_samples, spec.dim), tuple(xp.shape)))
      self._keep.append(xp)
      m.x0_paths_dev = xp.data_ptr()
    extra = getattr(spec, 'device_arrays', None)
    if extra is not None:
      mat, vec = extra(self.dtype)
      self._keep += [mat, vec]
      m.matrix, m.vector = mat.ctypes.data, vec.ctypes.data

    r = _lib.RngDesc()
    r.type, r.antithetic, r.skip = rng.type, int(rng.antithetic), rng.skip
    r.unit_stride, r.unit_offset = rng.unit_stride, rng.unit_offset
    if rng.type == _lib.RNG_PHILOX:
      r.key = rng.key
      r.counter = rng.counter
    elif rng.type == _lib.RNG_SOBOL:
      dn = sobol.direction_numbers(self.num_steps_total * spec.num_factors)
      self._keep.append(dn)
      r.direction_numbers = dn.ctypes.data
    else:
      draws = _tensor.from_dlpack(rng.normal_draws)
      want = (self.num_samples, self.num_steps_total, spec.num_factors)
      if tuple(draws.shape) != want:
        raise ValueError('normal_draws must have shape {} but has {}'.format(
            want, tuple(draws.shape)))
      draws = draws.to(device=_tensor.device(),
                       dtype=_tensor.torch_dtype(self.dtype)).contiguous()
      self._keep.append(draws)
      r.draws_dev = draws.data_ptr()

    _lib.require_cuda()
    handle = C.c_void_p()
    _lib.check(_lib.lib().tqf_plan_create(C.byref(m), C.byref(r),
                                          self.num_samples, C.byref(handle)))
    self._handle = handle

  def release(self):
    """`close()` unless the plan lives in the plan cache (`cached_plan`)."""
    if not self.cached:
      self.close()

  def close(self):
    if getattr(self, '_handle', None):
      _lib.lib().tqf_plan_destroy(self._handle)
      self._handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  def paths(self, record_slot, num_times, unit_offset=0, unit_count=None,
            exp_transform=False, out=None, column_sums=False):
    """States at the recorded steps: a `[rows, num_times, dim]` VIEW of a
    time-major `[num_times, dim, rows]` buffer (coalesced stores, no
    transpose).  rows = units (x2 for antithetic: partners follow).
    `exp_transform` stores exp(state) (log-space models).  `column_sums=True`
    returns `(paths, sums)` with `sums` a float64 device tensor
    `[num_times, dim]`, the sum over the rows of every stored value
    (accumulated by the kernel that writes them); `least_square_mc(...,
    column_sums=sums)` then skips its own pass over the paths."""
    unit_count = self.units - unit_offset if unit_count is None else unit_count
    rows = unit_count * (2 if self.rng.antithetic else 1)
    dim = self.spec.dim
    buf = _tensor.empty((num_times, dim, rows), self.dtype) if out is None else out
    assert tuple(buf.shape) == (num_times, dim, rows) and buf.is_contiguous()
    rec = np.ascontiguousarray(record_slot, dtype=np.int32)
    if column_sums:
      sums = torch.empty((num_times, dim), dtype=torch.float64, device=buf.device)
      _lib.check(_lib.lib().tqf_plan_paths_sums(
          self._handle, unit_offset, unit_count, rec.ctypes.data, buf.data_ptr(),
          1, dim * rows, rows,
          _lib.TRANSFORM_EXP if exp_transform else _lib.TRANSFORM_NONE,
          num_times, sums.data_ptr(), _tensor.current_stream_ptr()))
      return buf.permute(2, 0, 1), sums
    _lib.check(_lib.lib().tqf_plan_paths(
        self._handle, unit_offset, unit_count, rec.ctypes.data, buf.data_ptr(),
        1, dim * rows, rows,
        _lib.TRANSFORM_EXP if exp_transform else _lib.TRANSFORM_NONE,
        _tensor.current_stream_ptr()))
    return buf.permute(2, 0, 1)

  def set_sobol_clamp(self, clamp=True):
    """float32 Sobol draws whose uniform rounds to exactly 1.0 (possible beyond
    2^24 points): strict mode (default) reproduces the reference's +inf
    normal and the path is counted as non-finite; the clamped mode uses the largest
    float32 below one instead (documented non-reference mode)."""
    _lib.check(_lib.lib().tqf_plan_set_sobol_clamp(self._handle, int(bool(clamp))))

  def set_peer_exchange(self, peer_exchange):
    """Several GPUs of one box: every `price_sums` then returns the sums of ALL
    ranks, added inside the reduction kernel over NVLink peer memory
    (`tff_b200.distributed.PeerExchange`) -- no NCCL call per pricing.  All ranks
    must issue the same sequence of `price_sums` calls."""
    self._peer_exchange = peer_exchange
    _lib.check(_lib.lib().tqf_plan_set_peer_exchange(
        self._handle, peer_exchange.rank, peer_exchange.world, peer_exchange.ptrs,
        peer_exchange.epoch))

  def clear_peer_exchange(self):
    """Back to single-GPU pricing (a cached plan may have been used sharded before)."""
    px = getattr(self, '_peer_exchange', None)
    if px is not None:
      own = (C.c_void_p * 1)(px.ptrs[px.rank])
      _lib.check(_lib.lib().tqf_plan_set_peer_exchange(self._handle, 0, 1, own, 0))
      self._peer_exchange = None

  def price_sums(self, payoffs, unit_offset=0, unit_count=None):
    """Unnormalised per-payoff sums as a device tensor [num_payoffs, 4]:
    (sum, sum of squares, number of non-finite payoffs, 0)."""
    px = getattr(self, '_peer_exchange', None)
    if px is not None:
      # the buffers may have been used by another plan / the LSM passes meanwhile
      _lib.check(_lib.lib().tqf_plan_set_peer_exchange(
          self._handle, px.rank, px.world, px.ptrs, px.epoch))
    unit_count = self.units - unit_offset if unit_count is None else unit_count
    descs = (_lib.PayoffDesc * len(payoffs))(*[p.desc() for p in payoffs])
    # (fully overwritten by the reduction kernel)
    sums = torch.empty((len(payoffs), 4), dtype=torch.float64,
                       device=_tensor.device())
    _lib.check(_lib.lib().tqf_plan_price(
        self._handle, unit_offset, unit_count, descs, len(payoffs),
        sums.data_ptr(), _tensor.current_stream_ptr()))
    if px is not None:
      # the library advances the epoch only once the exchange kernel is launched: a
      # call that failed validation leaves every rank's count untouched
      ep = C.c_uint64()
      _lib.check(_lib.lib().tqf_plan_peer_epoch(self._handle, C.byref(ep)))
      px.epoch = int(ep.value)
    return sums


class HostPricing:
  """A plan with its payoff descriptors and result buffers bound once: `sums()` is a
  single FFI call (`tqf_plan_price_host`: kernels, read-back, synchronisation).  What a
  repeated pricing call with identical arguments runs (`euler_sampling.price`)."""

  def __init__(self, plan, payoffs):
    self.plan = plan
    self.count = len(payoffs)
    self.descs = (_lib.PayoffDesc * self.count)(*[p.desc() for p in payoffs])
    self.sums_dev = torch.empty((self.count, 4), dtype=torch.float64, device=_tensor.device())
    self.sums_host = np.empty((self.count, 4), dtype=np.float64)
    self._fn = _lib.lib().tqf_plan_price_host
    self._args = (plan._handle, 0, plan.units, self.descs, self.count, self.sums_dev.data_ptr(),
                  self.sums_host.ctypes.data)

  def sums(self):
    plan = self.plan
    if getattr(plan, '_peer_exchange', None) is not None:
      plan.clear_peer_exchange()
    _lib.check(self._fn(*self._args, torch.cuda.current_stream().cuda_stream))
    return self.sums_host

  def alive(self):
    return getattr(self.plan, '_handle', None) is not None


# ------------------------------------------------------------ plan cache ----
# Repeated pricing calls with identical inputs (a calibration loop re-pricing on the
# same grid, a benchmark) re-use the device-resident tables instead of allocating,
# uploading and freeing them per call: plans are keyed by the CONTENT of everything
# that goes to the device ("copied by the library into its own device
# cache keyed by content hash").  Least recently used plans are destroyed.
_PLAN_CACHE = {}
_PLAN_CACHE_SIZE = 8


def cached_plan(spec, all_times, num_steps, x0, rng, num_samples, dtype):
  """A `Plan` for these inputs from the cache, built on a miss.  Callers use
  `plan.release()` instead of `close()`.  Plans that reference caller-owned device
  memory (per-path initial states, `normal_draws=`) are never cached."""
  dtype = np.dtype(dtype)
  if rng.type == _lib.RNG_DRAWS:
    return Plan(spec, all_times, num_steps, x0, rng, num_samples, dtype)
  all_times = np.asarray(all_times, dtype=dtype)
  table = np.ascontiguousarray(spec.coef_table(all_times, dtype)[:int(num_steps)], dtype=np.float64)
  x0 = np.ascontiguousarray(np.asarray(x0, dtype=dtype).reshape(-1), dtype=np.float64)
  extra = getattr(spec, 'device_arrays', None)
  extra_bytes = b''.join(a.tobytes() for a in extra(dtype)) if extra is not None else b''
  rkey = (tuple(rng.key), tuple(rng.counter)) if rng.type == _lib.RNG_PHILOX else ()
  key = (torch.cuda.current_device() if torch.cuda.is_available() else -1,
         spec.kind, spec.dim, spec.num_factors, int(num_steps), all_times.shape[0] - 1, dtype.str,
         int(num_samples), rng.type, rng.antithetic, rkey, rng.skip, rng.unit_stride,
         rng.unit_offset, int(getattr(spec, 'exact_log', False)), table.tobytes(), x0.tobytes(),
         extra_bytes)
  plan = _PLAN_CACHE.pop(key, None)
  if plan is None or getattr(plan, '_handle', None) is None:
    plan = Plan(spec, all_times, num_steps, x0, rng, num_samples, dtype, table=table)
    plan.cached = True
  _PLAN_CACHE[key] = plan                 # most recently used last
  while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
    old = _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    old.close()
  return plan


def clear_plan_cache():
  while _PLAN_CACHE:
    _PLAN_CACHE.popitem()[1].close()


def measure_fma_peaks():
  """(DFMA/s, FFMA/s) issue peaks of the current device."""
  _lib.require_cuda()
  d, f = C.c_double(), C.c_double()
  _lib.check(_lib.lib().tqf_measure_fp64_peak(C.byref(d), C.byref(f)))
  return d.value, f.value
