"""Drift / volatility callables that the device engine can recognise.

The reference accepts arbitrary Python `drift_fn(t, x)` / `volatility_fn(t, x)`
(`models/euler_sampling.py:27-47`).  A CUDA kernel cannot run Python, so the
engine recognises callables that carry a `ModelSpec` (`tqf_spec`): the ones the
model classes of this package return, or ones built with the helpers below.
They stay ordinary callables -- `fn(t, x)` evaluates on the host with torch --
so code that inspects drift / volatility values keeps working.  Anything else
is rejected with NotImplementedError; there is no silent CPU fallback.
"""
import numpy as np
import torch

from tff_b200 import engine


class DeviceClosure:
  """A `drift_fn` or `volatility_fn` bound to a device model spec."""

  def __init__(self, spec, role, host_fn):
    self.tqf_spec = spec
    self.role = role
    self._host_fn = host_fn

  def __call__(self, t, x):
    return self._host_fn(t, x)


def resolve_spec(drift_fn, volatility_fn, dim=None):
  """The ModelSpec shared by a (drift_fn, volatility_fn) pair."""
  ds = getattr(drift_fn, 'tqf_spec', None)
  vs = getattr(volatility_fn, 'tqf_spec', None)
  if ds is None and vs is None and dim is not None and callable(drift_fn) and callable(
      volatility_fn):
    # plain Python callables: accepted when they are affine in the state
    return engine.ProbedAffineSpec(dim, drift_fn, volatility_fn)
  if ds is None or vs is None:
    raise NotImplementedError(
        'The B200 path engine cannot run arbitrary Python drift/volatility '
        'callables inside a CUDA kernel. Pass the closures of a model class '
        '(GeometricBrownianMotion, HestonModel, HullWhiteModel1F, '
        'MultivariateGeometricBrownianMotion) or build them with '
        'tff_b200.models.closures.affine_closures / gbm_closures. '
        'There is no CPU fallback.')
  if ds is not vs:
    raise NotImplementedError(
        '`drift_fn` and `volatility_fn` must come from the same device model')
  return ds


def _as_tensor(x, like=None):
  if isinstance(x, torch.Tensor):
    return x
  x = np.asarray(x)
  t = torch.as_tensor(x)
  if like is not None:
    t = t.to(device=like.device, dtype=like.dtype)
  return t


def _p(param, t, like):
  """Parameter value at scalar time `t` as a tensor like `like`."""
  if callable(param):
    tt = float(t.item() if isinstance(t, torch.Tensor) else t)
    v = np.asarray(param(np.asarray([tt])))[0]
  else:
    v = np.asarray(param)
  return torch.as_tensor(v, dtype=like.dtype, device=like.device)


def affine_closures(a0, a1, b, b1=0.0):
  """(drift_fn, volatility_fn) of dX = (a0(t) + a1(t) X) dt + (b(t) + b1(t) X) dW, dim 1.

  `a0`, `a1`, `b` are scalars or callables of an array of times.  Covers the
  log-space GBM of the reference's Monte-Carlo notebook
  (`examples/jupyter_notebooks/Monte_Carlo_Euler_Scheme.ipynb:223-275`).
  """
  spec = engine.AffineSpec1F(a0, a1, b, b1)

  def drift(t, x):
    x = _as_tensor(x)
    return _p(a0, t, x) + _p(a1, t, x) * x

  def vol(t, x):
    x = _as_tensor(x)
    return (_p(b, t, x) + _p(b1, t, x) * x).unsqueeze(-1)
  return DeviceClosure(spec, 'drift', drift), DeviceClosure(spec, 'volatility', vol)


def affine_tangent_closures(a0, a1, b, b1=0.0, da0=0.0, da1=0.0, db=0.0, db1=0.0):
  """`affine_closures` whose sampler also carries the pathwise tangents
  `dX/dX0` and `dX/dtheta` (the reference's `watch_params` route,
  `euler_sampling.py:393-402`): `sample` then returns `[N, k, 3]` with the
  components `[X, dX/dX0, dX/dtheta]`, `price` accepts the `*_tangent` payoffs.
  `da0 .. db1` are the derivatives of `a0 .. b1` with respect to the watched
  scalar parameter `theta`.  Log-space GBM of the Monte-Carlo notebook
  (`Monte_Carlo_Euler_Scheme.ipynb` cells 22-28), theta = sigma:
  `affine_tangent_closures(r - sigma**2 / 2, 0, sigma, da0=-sigma, db=1)`."""
  spec = engine.TangentAffineSpec1F(a0, a1, b, b1, da0, da1, db, db1)

  def drift(t, x):
    x = _as_tensor(x)
    return _p(a0, t, x) + _p(a1, t, x) * x

  def vol(t, x):
    x = _as_tensor(x)
    return (_p(b, t, x) + _p(b1, t, x) * x).unsqueeze(-1)
  return DeviceClosure(spec, 'drift', drift), DeviceClosure(spec, 'volatility', vol)


def gbm_closures(mean, volatility):
  """(drift_fn, volatility_fn) of dX = mean(t) X dt + volatility(t) X dW."""
  spec = engine.GbmSpec1F(mean, volatility)

  def drift(t, x):
    x = _as_tensor(x)
    return _p(mean, t, x) * x

  def vol(t, x):
    x = _as_tensor(x)
    return _p(volatility, t, x) * x.unsqueeze(-1)
  return DeviceClosure(spec, 'drift', drift), DeviceClosure(spec, 'volatility', vol)


def heston_closures(mean_reversion, theta, volvol, rho):
  """Heston closures (`heston/heston_model.py:143-173`)."""
  spec = engine.HestonEulerSpec(mean_reversion, theta, volvol, rho)

  def vol(t, x):
    x = _as_tensor(x)
    v = torch.sqrt(torch.abs(x[..., 1]))
    zeros = torch.zeros_like(v)
    r, vv = _p(rho, t, x), _p(volvol, t, x)
    col1 = torch.stack([v, vv * r * v], -1)
    col2 = torch.stack([zeros, vv * torch.sqrt(1 - r**2) * v], -1)
    return torch.stack([col1, col2], -1)

  def drift(t, x):
    x = _as_tensor(x)
    var = x[..., 1]
    k, th = _p(mean_reversion, t, x), _p(theta, t, x)
    return torch.stack([-var / 2, k * (th - var)], -1)
  return DeviceClosure(spec, 'drift', drift), DeviceClosure(spec, 'volatility', vol)


def heston_tangent_closures(mean_reversion, theta, volvol, rho, d_mean_reversion=0.0, d_theta=0.0,
                            d_volvol=0.0, d_rho=0.0, d_initial_state=(0.0, 0.0)):
  """`heston_closures` whose sampler also carries the pathwise tangents of `(X, V)` with
  respect to one scalar `p`: `sample` returns `[N, k, 4]` with the components
  `[X, V, dX/dp, dV/dp]`, `price` accepts the `*_tangent` payoffs on component
  `TangentHestonSpec.D_X`.  `d_*` are the derivatives of the parameters / of the initial
  state with respect to `p`: vega to the initial variance is `d_initial_state=(0, 1)`,
  the sensitivity to the vol-of-vol `d_volvol=1`, delta `d_initial_state=(1, 0)`."""
  spec = engine.TangentHestonSpec(mean_reversion, theta, volvol, rho, d_mean_reversion, d_theta,
                                  d_volvol, d_rho, d_initial_state)
  drift, vol = heston_closures(mean_reversion, theta, volvol, rho)
  return (DeviceClosure(spec, 'drift', drift._host_fn),   # pylint: disable=protected-access
          DeviceClosure(spec, 'volatility', vol._host_fn))   # pylint: disable=protected-access


def mvgbm_closures(means, volatilities, corr_matrix, dim):
  """(drift_fn, volatility_fn) of the correlated multi-asset GBM
  (`multivariate_geometric_brownian_motion.py:130-151`)."""
  spec = engine.MvGbmSpec(means, volatilities, corr_matrix, dim)

  def drift(t, x):
    del t
    x = _as_tensor(x)
    return torch.as_tensor(np.asarray(means), dtype=x.dtype, device=x.device) * x

  def vol(t, x):
    del t
    x = _as_tensor(x)
    vols = torch.as_tensor(np.asarray(volatilities), dtype=x.dtype, device=x.device) * x
    if corr_matrix is None:
      return torch.diag_embed(vols)
    chol = torch.linalg.cholesky(torch.as_tensor(np.asarray(corr_matrix), dtype=x.dtype,
                                                 device=x.device))
    return vols.unsqueeze(-1) * chol
  return DeviceClosure(spec, 'drift', drift), DeviceClosure(spec, 'volatility', vol)
