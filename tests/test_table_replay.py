"""Per-step coefficient tables replayed on the CPU against the oracle's Euler sampler.

The host half of every fused kernel is a `ModelSpec.coef_table`: one row per Euler
step holding `dt`, `sqrt(dt)` and the model parameters taken at `t_{i+1}`
(`euler_sampling.py:519`), sometimes pre-multiplied.  The device half is the model's
`step()` in `csrc/tqf_paths_kernel.cuh`.  Here each `step()` is written out in numpy,
line for line, and run over the table with the oracle's draws; the result must equal the
oracle's restatement of `euler_sampling.sample` with the corresponding closures -- so a
wrong column order, a parameter taken at the wrong end of a step, a missing `sqrt(dt)`
or a mis-formed product in a table shows up without a GPU.  The kernels themselves are
compared with the same oracle on the device (`tests/test_gpu_parity.py`,
`tests/test_tangents.py`, `tests/test_milstein.py`).
"""
import numpy as np
import pytest
import torch

from oracle import draws as odraws
from oracle import euler as oeuler
from oracle import grid as ogrid
from oracle import milstein as omilstein
from oracle import models as omodels
from oracle import tangent as otangent
from tff_b200 import engine
from tff_b200.math import piecewise

RT = odraws.RandomType
TIMES = np.array([0.13, 0.5, 0.77, 1.0])
TIME_STEP = 0.06          # uneven steps once the requested times are merged in
N = 64
SEED = [4, 2]


def _pw(jumps, values):
  """The same piecewise-constant parameter for the mirror and for the oracle."""
  return (piecewise.PiecewiseConstantFunc(jumps, values, dtype=np.float64),
          omodels.PiecewiseConstantFunc(jumps, values, dtype=np.float64))


def _replay(spec, step, x0, num_factors, dtype=np.float64, draw_columns=None):
  """Run `step(x, z, c)` over `spec.coef_table` with the oracle's draws; the recorded
  states `[N, len(TIMES), dim]` like the kernel's record plan (`keep_mask`).  `draw_columns`:
  width of the draw tensor the reference generates when it exceeds `num_factors`."""
  all_times, keep_mask, _ = ogrid.euler_grid(TIMES, dtype=dtype, time_step=TIME_STEP)
  table = spec.coef_table(all_times, dtype)
  steps = all_times.shape[0] - 1
  assert table.shape == (steps, spec.num_coef) and table.dtype == np.float64
  z = odraws.generate_mc_normal_draws(num_normal_draws=draw_columns or num_factors, num_time_steps=steps,
                                      num_sample_paths=N, batch_shape=(), random_type=RT.STATELESS, dtype=dtype,
                                      seed=SEED)[..., :num_factors]
  x = np.broadcast_to(np.asarray(x0, dtype=dtype), (N, len(x0))).copy()
  out = []
  for i in range(steps):
    x = step(x, z[i], table[i])
    if keep_mask[i + 1]:
      out.append(x.copy())
  return np.stack(out, axis=1)


def _oracle(dim, drift, vol, x0, **kw):
  return oeuler.sample(dim, drift, vol, TIMES, time_step=TIME_STEP, num_samples=N, initial_state=np.asarray(x0),
                       random_type=RT.STATELESS, seed=SEED, dtype=np.float64, **kw)


# ---- the kernels' step() functions, csrc/tqf_paths_kernel.cuh ---------------------------
def _affine_1f(x, z, c):            # AffineModel1F::step
  dw = z[:, 0] * c[1]
  dt_inc = c[0] * (c[2] + c[3] * x[:, 0])
  dw_inc = (c[4] + c[5] * x[:, 0]) * dw
  return ((x[:, 0] + dt_inc) + dw_inc)[:, None]


def _linear_1f(x, z, c):            # LinearModel1F::step
  return ((c[2] * x[:, 0] + c[3]) + c[4] * z[:, 0])[:, None]


def _gbm_1f(x, z, c):               # GbmModel1F::step
  dw = z[:, 0] * c[1]
  return ((x[:, 0] + c[0] * (c[2] * x[:, 0])) + (c[3] * x[:, 0]) * dw)[:, None]


def _milstein_1f(x, z, c):          # MilsteinAffine1FModel::step
  dw = z[:, 0] * c[1]
  vol = c[4] + c[5] * x[:, 0]
  hot = ((vol * c[5]) * (dw * dw - c[0])) / 2
  return (((x[:, 0] + c[0] * (c[2] + c[3] * x[:, 0])) + vol * dw) + hot)[:, None]


def _heston(x, z, c):               # HestonEulerModel::step
  var = x[:, 1]
  vol = np.sqrt(np.abs(var))
  return np.stack([vol * (z[:, 0] * c[0]) + (c[1] * var + x[:, 0]),
                   vol * (c[5] * z[:, 1] + c[4] * z[:, 0]) + (c[2] * (c[3] - var) + var)], -1)


def _affine_nd(d):
  def step(x, z, c):                # AffineModelND<D>::step
    dw = z * c[1]
    a0, a1, b = c[2:2 + d], c[2 + d:2 + d + d * d].reshape(d, d), c[2 + d + d * d:].reshape(d, d)
    return (x + c[0] * (a0 + x @ a1.T)) + dw @ b.T
  return step


def _tangent_affine(x, z, c):       # TangentAffine1FModel::step
  dw = z[:, 0] * c[1]
  xs = x[:, 0]
  g = c[5] * dw + c[0] * c[3]
  return np.stack([(xs + c[0] * (c[2] + c[3] * xs)) + (c[4] + c[5] * xs) * dw,
                   x[:, 1] * g + x[:, 1],
                   (x[:, 2] * g + x[:, 2]) + (c[0] * (c[6] + c[7] * xs) + (c[8] + c[9] * xs) * dw)], -1)


def _tangent_heston(x, z, c):       # TangentHestonModel::step
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
                   (vt + c[0] * (c[7] * (c[3] - v) + c[2] * (c[8] - vt))) + ((c[9] * s + c[4] * ds) * w + (c[4] * s) * wp)],
                  -1)


# ---- the cases -------------------------------------------------------------------------
def test_affine_1f_table():
  (a0, oa0), (a1, oa1) = _pw([0.4], [0.03, -0.02]), _pw([0.2, 0.8], [-0.5, -0.7, -0.1])
  (b0, ob0), (b1, ob1) = _pw([0.6], [0.2, 0.3]), _pw([0.5], [0.1, 0.05])
  spec = engine.AffineSpec1F(a0, a1, b0, b1)
  assert spec.kind == engine._lib.MODEL_AFFINE_1F
  got = _replay(spec, _affine_1f, [0.7], 1)
  p = lambda f, t: f(np.asarray([t]))[0]
  want = _oracle(1, lambda t, x: p(oa0, t) + p(oa1, t) * x, lambda t, x: (p(ob0, t) + p(ob1, t) * x)[..., None], [0.7])
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_additive_noise_runs_as_the_linear_model():
  # log-space GBM of C1 / C5: B = dt a0 and C = b0 sqrt(dt) are formed on the host
  (a0, oa0), (b0, ob0) = _pw([0.4], [0.03, -0.02]), _pw([0.6], [0.2, 0.3])
  spec = engine.AffineSpec1F(a0, 0.0, b0, 0.0)
  assert spec.kind == engine._lib.MODEL_LINEAR_1F and spec.num_coef == 5
  got = _replay(spec, _linear_1f, [0.7], 1)
  p = lambda f, t: f(np.asarray([t]))[0]
  want = _oracle(1, lambda t, x: p(oa0, t) + 0 * x, lambda t, x: (p(ob0, t) + 0 * x)[..., None], [0.7])
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_gbm_1f_table():
  (mu, omu), (sigma, osigma) = _pw([0.3], [0.05, 0.02]), _pw([0.1, 0.7], [0.2, 0.4, 0.1])
  got = _replay(engine.GbmSpec1F(mu, sigma), _gbm_1f, [100.0], 1)
  drift, vol = omodels.gbm_closures(omu, osigma, np.float64)
  np.testing.assert_allclose(got, _oracle(1, drift, vol, [100.0]), rtol=1e-12)


def test_heston_euler_table():
  (kappa, okappa), (theta, otheta) = _pw([0.5], [1.0, 1.1]), _pw([0.5], [0.04, 0.09])
  (xi, oxi), (rho, orho) = _pw([0.3], [0.5, 0.8]), _pw([0.5], [-0.7, 0.6])
  x0 = [np.log(100.0), 0.04]
  got = _replay(engine.HestonEulerSpec(kappa, theta, xi, rho), _heston, x0, 2)
  drift, vol = omodels.heston_closures(okappa, otheta, oxi, orho, np.float64)
  want = _oracle(2, drift, vol, x0)
  assert (want[..., 1] < 0).any()          # the |V| branch is exercised
  np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('d', [2, 3, 4])
def test_probed_affine_nd_table(d):
  rs = np.random.RandomState(d)
  a0, a1, b = rs.uniform(-0.2, 0.2, d), rs.uniform(-0.5, 0.5, (d, d)), rs.uniform(-0.3, 0.3, (d, d))
  # plain Python callables, as a user of `euler_sampling.sample` writes them (torch in, torch out)
  drift_t = lambda t, x: (1.0 + t) * torch.as_tensor(a0) + x @ torch.as_tensor(a1).T
  vol_t = lambda t, x: (torch.as_tensor(b) * torch.sqrt(0.5 + t)).expand(x.shape[0], d, d)
  spec = engine.ProbedAffineSpec(d, drift_t, vol_t)
  x0 = rs.uniform(-1.0, 1.0, d)
  spec.initial_state_hint = x0
  got = _replay(spec, _affine_nd(d), x0, d)
  want = _oracle(d, lambda t, x: (1.0 + t) * a0 + x @ a1.T,
                 lambda t, x: np.broadcast_to(b * np.sqrt(0.5 + t), x.shape[:-1] + (d, d)), x0)
  np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)      # the table is SOLVED from probes


def test_milstein_1f_table():
  (mu, omu), (sigma, osigma) = _pw([0.3], [0.05, 0.02]), _pw([0.1, 0.7], [0.2, 0.4, 0.1])
  spec = engine.MilsteinSpec1F(engine.GbmSpec1F(mu, sigma))
  # the reference draws dim + 3 dim order normals per step (order 5) and steps with the first
  # `dim` columns (`milstein_sampling.py:282-300`)
  got = _replay(spec, _milstein_1f, [100.0], 1, draw_columns=16)
  p = lambda f, t: f(np.asarray([t]))[0]
  want = omilstein.sample(dim=1, drift_fn=lambda t, x: p(omu, t) * x,
                          volatility_fn=lambda t, x: (p(osigma, t) * x)[..., None],
                          grad_volatility_fn=lambda t, x: p(osigma, t) * np.ones(x.shape + (1,)),
                          times=TIMES, time_step=TIME_STEP, num_samples=N, initial_state=np.array([100.0]),
                          random_type=RT.STATELESS, seed=SEED, dtype=np.float64)
  np.testing.assert_allclose(got, want, rtol=1e-12)


def test_tangent_affine_table():
  args = [_pw([0.4], [0.03, -0.02]), _pw([0.2], [-0.5, -0.7]), _pw([0.6], [0.2, 0.3]), _pw([0.5], [0.1, 0.05]),
          _pw([0.4], [1.0, 0.5]), _pw([0.3], [0.2, -0.1]), _pw([0.6], [0.7, 1.0]), _pw([0.5], [-0.3, 0.4])]
  spec = engine.TangentAffineSpec1F(*[a[0] for a in args])
  got = _replay(spec, _tangent_affine, spec.extend_initial_state([0.7]), 1)
  want = otangent.sample_with_tangents(*[a[1] for a in args], TIMES, [0.7], N, random_type=RT.STATELESS, seed=SEED,
                                       time_step=TIME_STEP)
  np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('which', range(4))
def test_tangent_heston_table(which):
  # one-hot parameter derivative (kappa, theta, xi, rho in turn) + derivative of the initial state
  p = [_pw([0.5], [1.0, 1.1]), _pw([0.5], [0.06, 0.09]), _pw([0.3], [0.3, 0.4]), _pw([0.5], [-0.7, 0.6])]
  d = [1.0 if i == which else 0.0 for i in range(4)]
  d0 = (0.0, 1.0) if which == 0 else (0.0, 0.0)
  spec = engine.TangentHestonSpec(*[q[0] for q in p], *d, d_initial_state=d0)
  x0 = [np.log(100.0), 0.08]
  got = _replay(spec, _tangent_heston, spec.extend_initial_state(np.asarray(x0)), 2)
  want = otangent.heston_with_tangents(*[q[1] for q in p], *d, d0, TIMES, x0, N, random_type=RT.STATELESS,
                                       seed=SEED, time_step=TIME_STEP)
  np.testing.assert_allclose(got[..., :2], want[..., :2], rtol=1e-11, atol=1e-13)
  np.testing.assert_allclose(got[..., 2:], want[..., 2:], rtol=1e-8, atol=1e-10)   # tangents carry 1 / (2 sqrt|V|)


# ---- Heston QE (`HestonModel.sample_paths`): its own grid (duplicates kept) and table ------
def _heston_qe(x, z, c):            # HestonQeModel::step / step_reference
  from scipy import special
  if c[0] == 0:                     # zero-length step: consumes its draws only
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
  return np.stack([xn, vn], -1)


@pytest.mark.parametrize('piecewise_params', [False, True])
def test_heston_qe_table(piecewise_params):
  from oracle import heston_qe as oqe
  from tff_b200.models.heston import qe
  if piecewise_params:
    # a jump ON a grid point gives a zero-length step (`heston_model.py:575-639` keeps duplicates)
    (kappa, okappa), (theta, otheta) = _pw([0.5], [1.0, 1.1]), _pw([0.5], [0.04, 0.09])
    (xi, oxi), (rho, orho) = _pw([0.3], [1.0, 0.8]), _pw([0.5], [-0.7, 0.6])
  else:
    kappa = okappa = 1.0
    theta = otheta = 0.04
    xi = oxi = 1.0                   # Feller violated: the exponential branch is taken
    rho = orho = -0.7
  x0 = np.array([np.log(100.0), 0.04])
  all_times, keep_mask = oqe.prepare_grid(TIMES, np.float64(0.05), np.dtype(np.float64), (okappa, otheta, oxi, orho))
  spec = qe.HestonQeSpec(kappa, theta, xi, rho, 1e-6)
  table = spec.coef_table(all_times, np.float64)
  steps = all_times.shape[0] - 1
  assert table.shape == (steps, 10)
  if piecewise_params:
    assert (table[:, 0] == 0).any()          # zero-length steps are marked inactive
  z = odraws.generate_mc_normal_draws(num_normal_draws=2, num_time_steps=steps, num_sample_paths=256, batch_shape=(),
                                      random_type=RT.STATELESS, dtype=np.float64, seed=SEED)
  x = np.broadcast_to(x0, (256, 2)).copy()
  out, took_exponential = [], False
  for i in range(steps):
    if table[i, 0] != 0:
      v = x[:, 1]
      m = table[i, 2] + (v - table[i, 2]) * table[i, 1]
      took_exponential |= bool(((v * table[i, 3] + table[i, 4]) / (m * m) >= 1.5).any())
    x = _heston_qe(x, z[i], table[i])
    if keep_mask[i + 1]:
      out.append(x.copy())
  got = np.stack(out, axis=1)
  want = oqe.sample_paths(okappa, otheta, oxi, orho, TIMES, x0, num_samples=256, random_type=RT.STATELESS, seed=SEED,
                          time_step=0.05)
  assert took_exponential
  np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


# ---- correlated multi-asset GBM: Cholesky factor and (mu, sigma) as kernel parameters ------
@pytest.mark.parametrize('dim', [3, 8])
def test_mvgbm_device_arrays(dim):
  rs = np.random.RandomState(dim)
  means, vols = rs.uniform(0.0, 0.1, dim), rs.uniform(0.1, 0.4, dim)
  a = rs.standard_normal((dim, dim))
  corr = a @ a.T
  corr = corr / np.sqrt(np.outer(np.diag(corr), np.diag(corr)))
  x0 = rs.uniform(50.0, 150.0, dim)
  spec = engine.MvGbmSpec(means, vols, corr, dim)
  chol, (mu, sg) = spec.device_arrays(np.float64)

  def step(x, z, c):                # mvgbm kernels, csrc/tqf_mvgbm.cu: x' = (x + dt mu x) + sigma x (L (z sqrt_dt))
    return (x + c[0] * (mu * x)) + (sg * x) * ((z * c[1]) @ chol.T)
  got = _replay(spec, step, x0, dim)
  drift, vol = omodels.mvgbm_closures(means, vols, corr, np.float64)
  np.testing.assert_allclose(got, _oracle(dim, drift, vol, x0), rtol=1e-12)

  # exact_log: the state is log x, the step adds (mu - sigma^2 / 2) dt + sqrt(dt) sigma (L z)
  # (`multivariate_geometric_brownian_motion.py:229-282`), sampled at the requested times only
  spec = engine.MvGbmSpec(means, vols, corr, dim, exact_log=True)
  chol, (mu_log, sg) = spec.device_arrays(np.float64)
  table = spec.coef_table(np.concatenate([[0.0], TIMES]), np.float64)
  z = odraws.generate_mc_normal_draws(num_normal_draws=dim, num_time_steps=len(TIMES), num_sample_paths=N,
                                      batch_shape=(), random_type=RT.STATELESS, dtype=np.float64, seed=SEED)
  x = np.broadcast_to(np.log(x0), (N, dim)).copy()
  out = []
  for i in range(len(TIMES)):
    x = x + (mu_log * table[i, 0] + table[i, 1] * sg * (z[i] @ chol.T))
    out.append(np.exp(x))
  want = omodels.mvgbm_exact_sample_paths(means, vols, corr, TIMES, initial_state=x0, num_samples=N,
                                          random_type=RT.STATELESS, seed=SEED)
  np.testing.assert_allclose(np.stack(out, axis=1), want, rtol=1e-12)
