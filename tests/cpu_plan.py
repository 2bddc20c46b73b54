"""TEST INFRASTRUCTURE: a numpy stand-in for `tff_b200.engine.Plan`.

The product has no CPU path (`_lib.require_cuda()` fails loudly without a GPU).  To run the
mirror's HOST logic -- argument handling, grids, batch layout, record plan, coefficient tables,
draw-unit bookkeeping, output layout -- in the `-m "not gpu"` suite, tests replace `engine.Plan`
by `CpuPlan` with pytest's `monkeypatch` (inside the test only).  `CpuPlan` keeps what the real
plan uploads and restates what the kernels do with it:

  * the model `step()` functions of `csrc/tqf_paths_kernel.cuh` / `csrc/tqf_mvgbm.cu`, line for
    line in numpy (`STEP`);
  * the draw addressing of `tqf.h`: path p of a plan uses draw unit `p * unit_stride +
    unit_offset`; unit u owns elements `[u D, (u + 1) D)` of the Philox stream, `D = steps *
    factors` (TensorFlow's `[units, D]` normal matrix, `models/utils.py:98-128`), or Sobol point
    `skip + 1 + u`; antithetic plans append the negated partners;
  * the record plan (`record_slot[s + 1]` = output slot of the state after step s);
  * the payoff evaluation of the fused mode (`eval_payoff` and the swaption payoff of
    `csrc/tqf_paths_kernel.cuh`: running extrema of state component 0 over the initial state and
    every executed step, the Brownian-bridge no-touch probabilities of `brownian_bridge=True`
    barriers, the claim evaluated at its `expiry_step`, sums / sums of squares / non-finite counts).

The streams themselves come from the oracle.  Nothing here is shipped or imported by the package.
"""
import numpy as np
import torch

from oracle import draws as odraws
from tff_b200 import _lib

RT = odraws.RandomType


# ---- the kernels' step() functions -------------------------------------------------------------
def affine_1f(x, z, c, spec=None):            # AffineModel1F::step
  dw = z[:, 0] * c[1]
  dt_inc = c[0] * (c[2] + c[3] * x[:, 0])
  dw_inc = (c[4] + c[5] * x[:, 0]) * dw
  return ((x[:, 0] + dt_inc) + dw_inc)[:, None]


def linear_1f(x, z, c, spec=None):            # LinearModel1F::step
  return ((c[2] * x[:, 0] + c[3]) + c[4] * z[:, 0])[:, None]


def gbm_1f(x, z, c, spec=None):               # GbmModel1F::step
  dw = z[:, 0] * c[1]
  return ((x[:, 0] + c[0] * (c[2] * x[:, 0])) + (c[3] * x[:, 0]) * dw)[:, None]


def milstein_1f(x, z, c, spec=None):          # MilsteinAffine1FModel::step
  dw = z[:, 0] * c[1]
  vol = c[4] + c[5] * x[:, 0]
  hot = ((vol * c[5]) * (dw * dw - c[0])) / 2
  return (((x[:, 0] + c[0] * (c[2] + c[3] * x[:, 0])) + vol * dw) + hot)[:, None]


def heston_euler(x, z, c, spec=None):         # HestonEulerModel::step
  var = x[:, 1]
  vol = np.sqrt(np.abs(var))
  return np.stack([vol * (z[:, 0] * c[0]) + (c[1] * var + x[:, 0]),
                   vol * (c[5] * z[:, 1] + c[4] * z[:, 0]) + (c[2] * (c[3] - var) + var)], -1)


def affine_nd(x, z, c, spec=None):            # AffineModelND<D>::step
  d = x.shape[1]
  dw = z * c[1]
  a0, a1, b = c[2:2 + d], c[2 + d:2 + d + d * d].reshape(d, d), c[2 + d + d * d:].reshape(d, d)
  return (x + c[0] * (a0 + x @ a1.T)) + dw @ b.T


def tangent_affine(x, z, c, spec=None):       # TangentAffine1FModel::step
  dw = z[:, 0] * c[1]
  xs = x[:, 0]
  g = c[5] * dw + c[0] * c[3]
  return np.stack([(xs + c[0] * (c[2] + c[3] * xs)) + (c[4] + c[5] * xs) * dw,
                   x[:, 1] * g + x[:, 1],
                   (x[:, 2] * g + x[:, 2]) + (c[0] * (c[6] + c[7] * xs) + (c[8] + c[9] * xs) * dw)], -1)


def tangent_heston(x, z, c, spec=None):       # TangentHestonModel::step
  dw0, dw1 = z[:, 0] * c[1], z[:, 1] * c[1]
  v, vt = x[:, 1], x[:, 3]
  s = np.sqrt(np.abs(v))
  with np.errstate(divide='ignore', invalid='ignore'):
    ds = np.where(s > 0, np.where(v < 0, -vt, vt) / (2 * s), 0.0)
  w = c[5] * dw0 + c[6] * dw1
  wp = c[10] * dw0 + c[11] * dw1
  return np.stack([(x[:, 0] + c[0] * (-0.5 * v)) + s * dw0,
                   (v + c[0] * (c[2] * (c[3] - v))) + (c[4] * s) * w,
                   (x[:, 2] + c[0] * (-0.5 * vt)) + ds * dw0,
                   (vt + c[0] * (c[7] * (c[3] - v) + c[2] * (c[8] - vt)))
                   + ((c[9] * s + c[4] * ds) * w + (c[4] * s) * wp)], -1)


def heston_qe(x, z, c, spec=None):            # HestonQeModel::step / step_reference
  from scipy import special
  if c[0] == 0:                               # zero-length step: consumes its draws only
    return x
  v = x[:, 1]
  m = c[2] + (v - c[2]) * c[1]
  s2 = v * c[3] + c[4]
  psi = s2 / (m * m)
  with np.errstate(all='ignore'):
    psi_inv = 2 / psi
    b2 = psi_inv - 1 + np.sqrt(psi_inv * (psi_inv - 1))
    quad = (m / (1 + b2)) * (np.sqrt(b2) + z[:, 0])**2
    p = (psi - 1) / (psi + 1)
    beta = (1 - p) / m
    u = 0.5 * (1 + special.erf(z[:, 0] * 0.70710678118654752440))
    expo = np.where(u > p, (np.log(1 - p) - np.log(1 - u)) / beta, 0.0)
  vn = np.where(psi < 1.5, quad, expo)
  xn = (((x[:, 0] + c[5]) + c[6] * v) + c[7] * vn) + np.sqrt(c[8] * v + c[9] * vn) * z[:, 1]
  return np.stack([xn, vn], -1).astype(x.dtype)


def hull_white_1f(x, z, c, spec=None):        # HullWhite1FModel::step, state [x, integral]
  xn = c[2] * z[:, 0] + (c[0] * x[:, 0] + c[1])
  return np.stack([xn, c[3] * xn + (x[:, 1] + c[4])], -1)


def hjm(x, z, c, spec):                       # HjmModel<F, NFS>::step, state [x_1..x_F, integral]
  f = x.shape[1] - 1
  dw = z[:, :f] * c[1]
  a0, kap = c[2:2 + f], c[2 + f:2 + 2 * f]
  b = c[2 + 2 * f:2 + 2 * f + f * f].reshape(f, f)
  w_pre, w_post, w_const = c[2 + 2 * f + f * f], c[3 + 2 * f + f * f], c[4 + 2 * f + f * f]
  xs = x[:, :f]
  xn = (xs + c[0] * (a0 - kap * xs)) + dw @ b.T
  integral = x[:, f] + (w_pre * xs.sum(axis=1) + w_post * xn.sum(axis=1) + w_const)
  return np.concatenate([xn, integral[:, None]], axis=1)


def mvgbm(x, z, c, spec):                     # csrc/tqf_mvgbm.cu
  chol, (mu, sg) = spec.device_arrays(x.dtype)
  chol, mu, sg = chol.astype(x.dtype), mu.astype(x.dtype), sg.astype(x.dtype)
  if getattr(spec, 'exact_log', False):       # state log x: exact log-normal increment
    return x + (mu * c[0] + c[1] * sg * (z @ chol.T))
  return (x + c[0] * (mu * x)) + (sg * x) * ((z * c[1]) @ chol.T)


STEP = {
    _lib.MODEL_AFFINE_1F: affine_1f, _lib.MODEL_LINEAR_1F: linear_1f, _lib.MODEL_GBM_1F: gbm_1f,
    _lib.MODEL_MILSTEIN_1F: milstein_1f, _lib.MODEL_HESTON_EULER: heston_euler, _lib.MODEL_AFFINE_ND: affine_nd,
    _lib.MODEL_AFFINE_1F_TANGENT: tangent_affine, _lib.MODEL_HESTON_TANGENT: tangent_heston,
    _lib.MODEL_HESTON_QE: heston_qe, _lib.MODEL_HW1F: hull_white_1f, _lib.MODEL_MVGBM: mvgbm,
    _lib.MODEL_HJM: hjm,
}


BRIDGE_VAR = {       # Model::bridge_var(pre-step state, coefficients): variance of the monitored increment
    _lib.MODEL_AFFINE_1F: lambda x, c: (c[4] + c[5] * x[:, 0])**2 * c[0],
    _lib.MODEL_LINEAR_1F: lambda x, c: np.full(x.shape[0], c[4] * c[4]),
    _lib.MODEL_HESTON_EULER: lambda x, c: np.abs(x[:, 1]) * (c[0] * c[0]),
}


def bridge_factor(level, xs, xe, var, upper):
  """`brownian_bridge_single`: P(no touch) = 1 - exp(-2 (x_s - b)(x_e - b) / var) when both ends are on the
  inner side of the barrier, 0 otherwise (`black_scholes/brownian_bridge.py:118-196`)."""
  ds, de = (level - xs, level - xe) if upper else (xs - level, xe - level)
  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    p = np.where(var > 0, 1.0 - np.exp(-2.0 * (ds * de) / np.where(var > 0, var, 1.0)), 1.0)
  return np.where((ds > 0) & (de > 0), p, 0.0)


def eval_payoff(d, x, xmax, xmin, surv_up=1.0, surv_dn=1.0):
  """`eval_payoff` / the swaption branch of the price kernel for all paths (float64)."""
  if d.kind == _lib.PAYOFF_HW_SWAPTION:
    nf = max(int(d.num_factors), 1)           # several factors (TQF_MODEL_HJM): log P_j = k_j - sum_i g_ji x_i
    acc = sum(d.pay_coef[j] * np.exp(d.pay_k[j] - sum(d.pay_g[j * nf + i] * x[:, i] for i in range(nf)))
              for j in range(d.num_payments))
    swap = np.exp(-x[:, -1]) * (1.0 - acc)
    return np.maximum(swap if d.is_payer else -swap, 0.0) * d.scale
  # component < 0: the basket mean over the assets (multi-asset kernels, csrc/tqf_mvgbm.cu)
  f, fmax, fmin = (x[:, d.component] if d.component >= 0 else x.mean(axis=1)), xmax, xmin
  tangent = x[:, d.tangent_component]
  fprime = 1.0
  with np.errstate(over='ignore', invalid='ignore'):
    if d.transform == _lib.TRANSFORM_EXP:
      f, fmax, fmin = np.exp(f), np.exp(fmax), np.exp(fmin)
      fprime = f
    call, put = f - d.strike, d.strike - f
    v = {
        _lib.PAYOFF_CALL: np.where(call > 0, call, 0.0),
        _lib.PAYOFF_PUT: np.where(put > 0, put, 0.0),
        _lib.PAYOFF_UP_OUT_CALL: np.where((call > 0) & ~(fmax > d.barrier), call, 0.0),
        _lib.PAYOFF_UP_OUT_PUT: np.where((put > 0) & ~(fmax > d.barrier), put, 0.0),
        _lib.PAYOFF_DOWN_OUT_PUT: np.where((put > 0) & ~(fmin < d.barrier), put, 0.0),
        _lib.PAYOFF_DOWN_OUT_CALL: np.where((call > 0) & ~(fmin < d.barrier), call, 0.0),
        _lib.PAYOFF_CALL_TANGENT: np.where(call > 0, fprime * tangent, 0.0),
        _lib.PAYOFF_PUT_TANGENT: np.where(put > 0, -(fprime * tangent), 0.0),
        _lib.PAYOFF_IDENTITY: f,
    }[d.kind]
    if d.brownian_bridge and d.kind in (_lib.PAYOFF_UP_OUT_CALL, _lib.PAYOFF_UP_OUT_PUT):
      v = v * surv_up
    if d.brownian_bridge and d.kind in (_lib.PAYOFF_DOWN_OUT_PUT, _lib.PAYOFF_DOWN_OUT_CALL):
      v = v * surv_dn
    return np.where(np.isfinite(f), v, np.nan) * d.scale


def unit_draws(rng, num_factors, steps_total, units, dtype):
  """Normals `[steps_total, units, num_factors]` of the plan's units (first halves for the
  antithetic types), addressed as `tqf.h` documents."""
  if rng.normal_draws is not None:
    z = rng.normal_draws.detach().cpu().numpy() if isinstance(rng.normal_draws, torch.Tensor) else np.asarray(
        rng.normal_draws)
    return np.transpose(z.astype(dtype), [1, 0, 2])
  base = {RT.PSEUDO_ANTITHETIC.value: RT.PSEUDO, RT.STATELESS_ANTITHETIC.value: RT.STATELESS}.get(
      rng.random_type.value, RT(rng.random_type.value))
  lo = rng.unit_offset
  hi = lo + (units - 1) * rng.unit_stride + 1
  rows = odraws._draws_of_path_range(num_factors, steps_total, 1 << 40, base, rng.skip, rng.seed, np.dtype(dtype),
                                     (lo, hi))
  return rows[:, ::rng.unit_stride]


class CpuPlan:
  """`engine.Plan` for the CPU suite: same constructor, `paths()`, `close()`, `release()`."""

  def __init__(self, spec, all_times, num_steps, x0, rng, num_samples, dtype, x0_paths=None, table=None):
    self.spec, self.rng, self.cached = spec, rng, False
    self.dtype = np.dtype(dtype)
    self.num_samples, self.num_steps = int(num_samples), int(num_steps)
    self.all_times = np.asarray(all_times, dtype=self.dtype)
    self.num_steps_total = self.all_times.shape[0] - 1
    if rng.antithetic and self.num_samples % 2 != 0:
      raise ValueError('First dimension of `sample_shape` should be even for PSEUDO_ANTITHETIC random type')
    self.units = self.num_samples // 2 if rng.antithetic else self.num_samples
    self.table = np.asarray(spec.coef_table(self.all_times, self.dtype) if table is None else table)[:self.num_steps]
    self.x0 = np.asarray(x0, dtype=self.dtype).reshape(-1)
    if self.x0.shape[0] != spec.dim:
      raise ValueError('initial state must have {} components'.format(spec.dim))
    self.x0_paths = None if x0_paths is None else np.asarray(x0_paths, dtype=self.dtype)

  def close(self):
    pass

  release = close

  def clear_peer_exchange(self):
    pass

  def _start(self, unit_offset=0, unit_count=None):
    """Draws and initial states of the units [unit_offset, unit_offset + unit_count) -- the range a
    rank owns in a sharded run; rows are [units | their antithetic partners]."""
    dt = self.dtype
    unit_count = self.units - unit_offset if unit_count is None else int(unit_count)
    assert 0 <= unit_offset and unit_offset + unit_count <= self.units
    z = unit_draws(self.rng, self.spec.num_factors, self.num_steps_total, self.units, dt)
    z = z[:, unit_offset:unit_offset + unit_count]
    sel = np.arange(unit_offset, unit_offset + unit_count)
    if self.rng.antithetic:
      z = np.concatenate([z, -z], axis=1)
      sel = np.concatenate([sel, sel + self.units])
    rows = z.shape[1]
    x = (np.broadcast_to(self.x0, (rows, self.spec.dim)) if self.x0_paths is None else self.x0_paths[sel]).astype(dt)
    return z, x, rows

  def price_sums(self, payoffs, unit_offset=0, unit_count=None):
    """`[num_payoffs, 4]`: sum, sum of squares, number of non-finite payoffs, 0 (`tqf_plan_price`)."""
    z, x, rows = self._start(unit_offset, unit_count)
    descs = [p.desc() for p in payoffs]
    barrier_kinds = (_lib.PAYOFF_UP_OUT_CALL, _lib.PAYOFF_UP_OUT_PUT, _lib.PAYOFF_DOWN_OUT_PUT, _lib.PAYOFF_DOWN_OUT_CALL)
    monitored = {d.component for d in descs if d.kind in barrier_kinds}
    assert len(monitored) <= 1                  # the kernel monitors ONE state component (tqf_paths.cu)
    mon = monitored.pop() if monitored else 0
    xmax, xmin = x[:, mon].astype(np.float64), x[:, mon].astype(np.float64)
    # continuous monitoring (tqf_paths.cu): one level per direction, in state space
    level = {}
    for d in descs:
      if d.kind in barrier_kinds and d.brownian_bridge:
        assert d.component == 0 and self.spec.kind in BRIDGE_VAR
        up = d.kind in (_lib.PAYOFF_UP_OUT_CALL, _lib.PAYOFF_UP_OUT_PUT)
        lv = np.log(d.barrier) if d.transform == _lib.TRANSFORM_EXP else d.barrier
        assert level.setdefault(up, lv) == lv
    surv_up, surv_dn = np.ones(rows), np.ones(rows)
    step = STEP[self.spec.kind]
    table = self.table.astype(self.dtype)
    out = np.zeros((len(descs), 4))
    for s in range(self.num_steps):
      if level:
        var = BRIDGE_VAR[self.spec.kind](x.astype(np.float64), table[s].astype(np.float64))
        pre = x[:, 0].astype(np.float64)
      x = np.asarray(step(x, z[s], table[s], self.spec), dtype=self.dtype)
      if True in level:
        surv_up = surv_up * bridge_factor(level[True], pre, x[:, 0].astype(np.float64), var, True)
      if False in level:
        surv_dn = surv_dn * bridge_factor(level[False], pre, x[:, 0].astype(np.float64), var, False)
      xmax, xmin = np.maximum(xmax, x[:, mon]), np.minimum(xmin, x[:, mon])
      for q, d in enumerate(descs):
        if (d.expiry_step if d.expiry_step > 0 else self.num_steps) == s + 1:
          v = eval_payoff(d, x.astype(np.float64), xmax, xmin, surv_up, surv_dn)
          ok = np.isfinite(v)
          out[q] = [v[ok].sum(), (v[ok]**2).sum(), float((~ok).sum()), 0.0]
    return torch.from_numpy(out)

  def paths(self, record_slot, num_times, unit_offset=0, unit_count=None, exp_transform=False, out=None,
            column_sums=False):
    dt = self.dtype
    z, x, rows = self._start(unit_offset, unit_count)
    buf = np.zeros((num_times, self.spec.dim, rows), dtype=dt)
    step = STEP[self.spec.kind]
    table = self.table.astype(dt)
    record_slot = np.asarray(record_slot)
    for s in range(-1, self.num_steps):
      if s >= 0:
        x = np.asarray(step(x, z[s], table[s], self.spec), dtype=dt)
      slot = record_slot[s + 1]
      if slot >= 0:
        buf[slot] = (np.exp(x) if exp_transform else x).T
    t = torch.from_numpy(buf)
    if out is not None:
      out.copy_(t)
      t = out
    if column_sums:        # float64 sums over the rows of every stored value (`tqf_plan_paths_sums`)
      return t.permute(2, 0, 1), torch.from_numpy(buf.astype(np.float64).sum(axis=2))
    return t.permute(2, 0, 1)


def install(monkeypatch):
  """Replace, for one test, everything of the mirror that touches the device: the plan and the
  stand-alone generator fills (`tff_b200.math.random.{philox, sobol, halton}`), whose outputs become
  CPU tensors holding the oracle's streams."""
  from oracle import halton as ohalton
  from oracle import philox as ophilox
  from oracle import sobol as osobol
  from tff_b200 import _tensor
  from tff_b200 import engine
  from tff_b200.math.random import halton, philox, sobol
  monkeypatch.setattr(engine, 'Plan', CpuPlan)
  monkeypatch.setattr(engine, 'cached_plan', lambda *a, **k: CpuPlan(*a, **k))
  monkeypatch.setattr(_tensor, 'device', lambda: torch.device('cpu'))
  monkeypatch.setattr(torch.cuda, 'current_device', lambda: 0)
  monkeypatch.setattr(philox, 'normal', lambda shape, dtype=None, seed=None: torch.from_numpy(
      ophilox.stateful_normal(tuple(shape), seed, np.dtype(dtype))))
  monkeypatch.setattr(philox, 'stateless_normal', lambda shape, seed, dtype=None: torch.from_numpy(
      ophilox.stateless_normal(tuple(shape), seed, np.dtype(dtype))))
  monkeypatch.setattr(sobol, 'sample_normal', lambda dim, n, skip=0, dtype=None: torch.from_numpy(
      odraws._erfinv_times_sqrt2(osobol.sample(dim, n, skip=skip, dtype=dtype), dtype)))

  def halton_normal(dim, n, skip=0, dtype=None, randomized=False, seed=None, randomization_params=None):
    assert randomization_params is None
    u = ohalton.sample(dim, sequence_indices=np.arange(skip, skip + n), dtype=dtype, randomized=randomized, seed=seed)
    return torch.from_numpy(odraws._erfinv_times_sqrt2(u, dtype))
  monkeypatch.setattr(halton, 'sample_normal', halton_normal)
