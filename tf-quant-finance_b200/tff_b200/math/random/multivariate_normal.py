"""`mv_normal_sample` on the device (`math/random_ops/multivariate_normal.py`)."""
import enum

import numpy as np
import torch

from tff_b200 import _tensor
from tff_b200.math.random import philox
from tff_b200.math.random import sobol


@enum.unique
class RandomType(enum.Enum):
  """Types of random number sequences (`multivariate_normal.py:27-44`)."""
  PSEUDO = 0
  STATELESS = 1
  HALTON = 2
  HALTON_RANDOMIZED = 3
  SOBOL = 4
  PSEUDO_ANTITHETIC = 5
  STATELESS_ANTITHETIC = 6


def _process_mean_scale(mean, scale_matrix, covariance_matrix, dtype):
  """`multivariate_normal.py:427-451`."""
  dev = _tensor.device()
  if scale_matrix is not None:
    dt = _tensor.infer_dtype(scale_matrix, dtype)
    scale_matrix = torch.as_tensor(_tensor.to_numpy(scale_matrix, dt), device=dev)
  elif covariance_matrix is not None:
    dt = _tensor.infer_dtype(covariance_matrix, dtype)
    cov = torch.as_tensor(_tensor.to_numpy(covariance_matrix, dt), device=dev)
    scale_matrix = torch.linalg.cholesky(cov)
  if mean is None:
    mean_t = None
    batch_shape = tuple(scale_matrix.shape[:-1])
    dt = _tensor.np_dtype(scale_matrix.dtype)
  else:
    dt = _tensor.infer_dtype(mean, dtype)
    mean_t = torch.as_tensor(_tensor.to_numpy(mean, dt), device=dev)
    batch_shape = tuple(mean_t.shape)
  return mean_t, scale_matrix, batch_shape, batch_shape[-1], dt


def _finish(samples, mean, scale_matrix):
  if scale_matrix is not None:
    samples = torch.matmul(scale_matrix, samples.unsqueeze(-1)).squeeze(-1)
  return samples if mean is None else mean + samples


def multivariate_normal(sample_shape, mean=None, covariance_matrix=None,
                        scale_matrix=None, random_type=None,
                        validate_args=False, seed=None, dtype=None, name=None,
                        **kwargs):
  """Draws from a multivariate normal; returns a CUDA tensor of shape
  `sample_shape + batch_shape` (`multivariate_normal.py:47-245`)."""
  del name, validate_args
  random_type = RandomType.PSEUDO if random_type is None else random_type
  if mean is None and covariance_matrix is None and scale_matrix is None:
    raise ValueError('At least one of mean, covariance_matrix or scale_matrix'
                     ' must be specified.')
  if covariance_matrix is not None and scale_matrix is not None:
    raise ValueError('Only one of covariance matrix or scale matrix'
                     ' must be specified')
  sample_shape = tuple(int(s) for s in np.asarray(
      _tensor.to_numpy(sample_shape)).reshape(-1))
  mean_t, scale_t, batch_shape, dim, dt = _process_mean_scale(
      mean, scale_matrix, covariance_matrix, dtype)
  rt = RandomType(random_type.value) if isinstance(random_type, enum.Enum) else random_type

  if rt in (RandomType.PSEUDO, RandomType.STATELESS,
            RandomType.PSEUDO_ANTITHETIC, RandomType.STATELESS_ANTITHETIC):
    anti = rt in (RandomType.PSEUDO_ANTITHETIC, RandomType.STATELESS_ANTITHETIC)
    shape = sample_shape
    if anti:
      if sample_shape[0] % 2 != 0:
        raise ValueError('First dimension of `sample_shape` should be even for '
                         'PSEUDO_ANTITHETIC random type')
      shape = (sample_shape[0] // 2,) + sample_shape[1:]
    if rt in (RandomType.PSEUDO, RandomType.PSEUDO_ANTITHETIC):
      raw = philox.normal(shape + batch_shape, dtype=dt, seed=seed)
    else:
      if seed is None:
        raise ValueError('`seed` should be specified if the `random_type` is '
                         '`STATELESS` or `STATELESS_ANTITHETIC`')
      raw = philox.stateless_normal(shape + batch_shape, seed, dtype=dt)
    result = _finish(raw, mean_t, scale_t)
    if not anti:
      return result
    if mean_t is None:
      return torch.cat([result, -result], dim=0)
    return torch.cat([result, 2 * mean_t - result], dim=0)

  if rt == RandomType.SOBOL:
    skip = int(kwargs.get('skip', 0) or 0)
    out_shape_t = tuple(reversed(batch_shape)) + sample_shape
    num_samples = int(np.prod(out_shape_t)) // dim
    z = sobol.sample_normal(dim, num_samples, skip=skip, dtype=dt)   # [n, dim]
    nb, ns = len(batch_shape), len(sample_shape)
    perm = list(range(nb, nb + ns)) + list(range(nb - 1, -1, -1))
    z = z.t().reshape(out_shape_t).permute(perm)
    return _finish(z, mean_t, scale_t)

  if rt in (RandomType.HALTON, RandomType.HALTON_RANDOMIZED):
    # multivariate_normal.py:391-424
    from tff_b200.math.random import halton  # pylint: disable=g-import-not-at-top
    skip = int(kwargs.get('skip', 0) or 0)
    out_shape_t = tuple(reversed(batch_shape)) + sample_shape
    num_samples = int(np.prod(out_shape_t)) // dim
    z = halton.sample_normal(dim, num_samples, skip=skip, dtype=dt,
                             randomized=rt == RandomType.HALTON_RANDOMIZED, seed=seed,
                             randomization_params=kwargs.get('randomization_params'))
    nb, ns = len(batch_shape), len(sample_shape)
    perm = list(range(nb, nb + ns)) + list(range(nb - 1, -1, -1))
    z = z.t().reshape(out_shape_t).permute(perm)
    return _finish(z, mean_t, scale_t)

  raise NotImplementedError(
      'Only STATELESS, PSEUDO, PSEUDO_ANTITHETIC, STATELESS_ANTITHETIC,  '
      'HALTON, HALTON_RANDOMIZED, and SOBOL random types are currently '
      'supported. Supplied: {}'.format(random_type))
