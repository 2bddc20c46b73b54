"""Oracle (test infrastructure): normal draws consumed by the samplers.

Restates
  * `math/random_ops/multivariate_normal.py:47-451` (`multivariate_normal`
    a.k.a. `mv_normal_sample`): mean-only inputs (the form every caller on the
    hot path uses) and the `covariance_matrix` / `scale_matrix` forms;
  * `models/utils.py:20-128` (`generate_mc_normal_draws`).
"""
import enum
import numpy as np
from scipy import special

from oracle import philox
from oracle import sobol

_SQRT_2 = np.sqrt(2.)


@enum.unique
class RandomType(enum.Enum):
  """`multivariate_normal.py:27-44`."""
  PSEUDO = 0
  STATELESS = 1
  HALTON = 2
  HALTON_RANDOMIZED = 3
  SOBOL = 4
  PSEUDO_ANTITHETIC = 5
  STATELESS_ANTITHETIC = 6


def _erfinv_times_sqrt2(u, dtype):
  """`tf.math.erfinv((u - 0.5) * 2) * sqrt(2)` (`multivariate_normal.py:420`).

  TF's erfinv(x) is ndtri(0.5 x + 0.5) / sqrt(2) (Eigen / Cephes); here the
  float64 Cephes `ndtri` of scipy is used and the result rounded to `dtype`.
  """
  dtype = np.dtype(dtype)
  x = (u - dtype.type(0.5)) * dtype.type(2)            # exact for dyadic u
  z = special.ndtri(0.5 * x.astype(np.float64) + 0.5)
  return z.astype(dtype)


def _process_mean_scale(mean, scale_matrix, covariance_matrix, dtype):
  """`multivariate_normal.py:427-451` -> (mean, scale, batch_shape, dtype)."""
  if scale_matrix is not None:
    scale_matrix = np.asarray(scale_matrix, dtype=dtype)
  elif covariance_matrix is not None:
    scale_matrix = np.linalg.cholesky(np.asarray(covariance_matrix, dtype=dtype))
  if mean is None:
    dtype = scale_matrix.dtype
    return dtype.type(0.0), scale_matrix, tuple(scale_matrix.shape[:-1]), dtype
  mean = np.asarray(mean, dtype=dtype)
  return mean, scale_matrix, tuple(mean.shape), mean.dtype


def _shift_scale(mean, scale_matrix, samples):
  """`mean + tf.linalg.matvec(scale_matrix, samples)` (`multivariate_normal.py:270-273`)."""
  if scale_matrix is None:
    return mean + samples
  return (mean + np.einsum('...ij,...j->...i', scale_matrix, samples)).astype(samples.dtype)


def _mvnormal_pseudo(sample_shape, mean, random_type, seed, dtype, scale_matrix=None,
                     batch_shape=None):
  """`multivariate_normal.py:248-273`."""
  out_shape = tuple(sample_shape) + tuple(mean.shape if batch_shape is None else batch_shape)
  if random_type == RandomType.PSEUDO:
    samples = philox.stateful_normal(out_shape, seed, dtype)
  else:
    if seed is None:
      raise ValueError('`seed` should be specified if the `random_type` is '
                       '`STATELESS` or `STATELESS_ANTITHETIC`')
    samples = philox.stateless_normal(out_shape, seed, dtype)
  return _shift_scale(mean, scale_matrix, samples)


def _mvnormal_pseudo_antithetic(sample_shape, mean, random_type, seed, dtype, scale_matrix=None,
                                batch_shape=None, mean_is_none=False):
  """`multivariate_normal.py:276-311`."""
  n0 = int(sample_shape[0])
  if n0 % 2 != 0:
    raise ValueError('First dimension of `sample_shape` should be even for '
                     'PSEUDO_ANTITHETIC random type')
  half_shape = (n0 // 2,) + tuple(sample_shape[1:])
  base = (RandomType.PSEUDO if random_type == RandomType.PSEUDO_ANTITHETIC
          else RandomType.STATELESS)
  r = _mvnormal_pseudo(half_shape, mean, base, seed, dtype, scale_matrix, batch_shape)
  if mean_is_none:
    return np.concatenate([r, -r], axis=0)
  return np.concatenate([r, 2 * mean - r], axis=0)


def _mvnormal_sobol(sample_shape, mean, skip, dtype, scale_matrix=None, batch_shape=None):
  """`multivariate_normal.py:356-424` for SOBOL."""
  batch_shape = tuple(mean.shape if batch_shape is None else batch_shape)
  dim = batch_shape[-1]
  sample_shape = tuple(int(s) for s in sample_shape)
  output_shape_t = tuple(reversed(batch_shape)) + sample_shape
  num_samples = int(np.prod(output_shape_t)) // dim
  seq = sobol.sample(dim, num_samples, skip=skip, dtype=dtype)   # [n, dim]
  seq = seq.T
  size_sample = len(sample_shape)
  size_batch = len(batch_shape)
  perm = (list(range(size_batch, size_batch + size_sample)) +
          list(range(size_batch - 1, -1, -1)))
  seq = np.transpose(seq.reshape(output_shape_t), perm)
  return _shift_scale(mean, scale_matrix, _erfinv_times_sqrt2(seq, dtype))


def _mvnormal_halton(sample_shape, mean, skip, dtype, randomized=False, seed=None,
                     scale_matrix=None, batch_shape=None):
  """`multivariate_normal.py:356-424` for HALTON / HALTON_RANDOMIZED:
  `halton.sample(dim, sequence_indices=range(skip, skip + n))`, then the same
  transpose / reshape / erfinv as the Sobol branch."""
  from oracle import halton  # pylint: disable=g-import-not-at-top
  batch_shape = tuple(mean.shape if batch_shape is None else batch_shape)
  dim = batch_shape[-1]
  sample_shape = tuple(int(s) for s in sample_shape)
  output_shape_t = tuple(reversed(batch_shape)) + sample_shape
  num_samples = int(np.prod(output_shape_t)) // dim
  seq = halton.sample(dim, sequence_indices=np.arange(skip, skip + num_samples), dtype=dtype,
                      randomized=randomized, seed=seed)
  seq = seq.T
  size_sample = len(sample_shape)
  size_batch = len(batch_shape)
  perm = (list(range(size_batch, size_batch + size_sample)) +
          list(range(size_batch - 1, -1, -1)))
  seq = np.transpose(seq.reshape(output_shape_t), perm)
  return _shift_scale(mean, scale_matrix, _erfinv_times_sqrt2(seq, dtype))


def mv_normal_sample(sample_shape, mean=None, random_type=None, seed=None,
                     dtype=None, skip=0, covariance_matrix=None, scale_matrix=None):
  """`multivariate_normal` (`multivariate_normal.py:47-245`)."""
  random_type = RandomType.PSEUDO if random_type is None else random_type
  random_type = RandomType(random_type.value)
  if mean is None and covariance_matrix is None and scale_matrix is None:
    raise ValueError('At least one of mean, covariance_matrix or scale_matrix must be specified.')
  if covariance_matrix is not None and scale_matrix is not None:
    raise ValueError('Only one of covariance matrix or scale matrix must be specified')
  mean_is_none = mean is None
  mean, scale, batch_shape, dtype = _process_mean_scale(mean, scale_matrix, covariance_matrix, dtype)
  if random_type in (RandomType.PSEUDO, RandomType.STATELESS):
    return _mvnormal_pseudo(sample_shape, mean, random_type, seed, dtype, scale, batch_shape)
  if random_type in (RandomType.PSEUDO_ANTITHETIC,
                     RandomType.STATELESS_ANTITHETIC):
    return _mvnormal_pseudo_antithetic(sample_shape, mean, random_type, seed,
                                       dtype, scale, batch_shape, mean_is_none)
  if random_type == RandomType.SOBOL:
    return _mvnormal_sobol(sample_shape, mean, skip, dtype, scale, batch_shape)
  if random_type == RandomType.HALTON:
    return _mvnormal_halton(sample_shape, mean, skip, dtype, scale_matrix=scale, batch_shape=batch_shape)
  if random_type == RandomType.HALTON_RANDOMIZED:
    return _mvnormal_halton(sample_shape, mean, skip, dtype, randomized=True, seed=seed,
                            scale_matrix=scale, batch_shape=batch_shape)
  raise NotImplementedError(
      'Only STATELESS, PSEUDO, PSEUDO_ANTITHETIC, STATELESS_ANTITHETIC and '
      'SOBOL are restated by the oracle. Supplied: {}'.format(random_type))


def uniform(dim, sample_shape, random_type=None, dtype=None, seed=None, skip=0):
  """`tff.math.random.uniform` (`math/random_ops/uniform.py:25-153`)."""
  random_type = RandomType.PSEUDO if random_type is None else RandomType(random_type.value)
  dtype = np.dtype(dtype or np.float32)
  sample_shape = [int(s) for s in sample_shape]
  shape = sample_shape + [int(dim)]
  if random_type == RandomType.PSEUDO:
    return philox.stateful_uniform(shape, seed, dtype)
  if random_type == RandomType.STATELESS:
    if seed is None:
      raise ValueError('`seed` must be supplied if the `random_type` is STATELESS.')
    return philox.stateless_uniform(shape, seed, dtype)
  if random_type == RandomType.PSEUDO_ANTITHETIC:
    raise NotImplementedError('At the moment antithetic sampling is not supported for the uniform '
                              'distribution.')
  num = int(np.prod(sample_shape))
  if random_type == RandomType.SOBOL:
    seq = sobol.sample(int(dim), num, skip=skip, dtype=dtype)
  else:   # uniform.py:135-150: everything else is a Halton sequence
    from oracle import halton  # pylint: disable=g-import-not-at-top
    seq = halton.sample(int(dim), sequence_indices=np.arange(skip, skip + num), dtype=dtype,
                        randomized=random_type == RandomType.HALTON_RANDOMIZED, seed=seed)
  return seq.reshape(shape)


def _draws_of_path_range(num_normal_draws, num_time_steps, num_sample_paths,
                         random_type, skip, seed, dtype, path_range):
  """Rows [lo, hi) of the `[num_sample_paths, steps * draws]` matrix that
  `generate_mc_normal_draws` builds, WITHOUT building the rest (chunked /
  multi-process oracle runs at the BASELINE sizes).  Row p of the Philox matrix
  is elements [p D, (p + 1) D) of the flat stream, row p of the Sobol matrix is
  point skip + 1 + p (its value does not depend on `num_digits`).  For the
  antithetic types [lo, hi) addresses the first half and the partners
  `2 mean - r = -r` follow, like `_mvnormal_pseudo_antithetic`."""
  lo, hi = int(path_range[0]), int(path_range[1])
  d = num_time_steps * num_normal_draws
  anti = random_type in (RandomType.PSEUDO_ANTITHETIC, RandomType.STATELESS_ANTITHETIC)
  limit = num_sample_paths // 2 if anti else num_sample_paths
  if not 0 <= lo <= hi <= limit:
    raise ValueError('path_range outside the run')
  if random_type == RandomType.SOBOL:
    u = sobol.sample(d, hi - lo, skip=skip + lo, dtype=dtype)
    rows = _erfinv_times_sqrt2(u, dtype)
  elif random_type in (RandomType.PSEUDO, RandomType.PSEUDO_ANTITHETIC):
    key, ctr = philox.stateful_key_counter(seed)
    rows = philox.normal_fill(key, ctr, (hi - lo) * d, dtype, first_element=lo * d)
  elif random_type in (RandomType.STATELESS, RandomType.STATELESS_ANTITHETIC):
    key, ctr = philox.stateless_key_counter(seed)
    rows = philox.normal_fill(key, ctr, (hi - lo) * d, dtype, first_element=lo * d)
  else:
    raise NotImplementedError(random_type)
  rows = rows.reshape(hi - lo, num_time_steps, num_normal_draws)
  if anti:
    rows = np.concatenate([rows, -rows], axis=0)
  return np.transpose(rows, [1, 0, 2])


def generate_mc_normal_draws(num_normal_draws, num_time_steps,
                             num_sample_paths, random_type, batch_shape=None,
                             skip=0, seed=None, dtype=None, path_range=None):
  """`models/utils.py:20-128` -> [steps] + batch_shape + [paths, draws].

  `path_range=(lo, hi)` (oracle extension, no batch): only the paths lo..hi-1 of
  the `num_sample_paths`-path call (for the antithetic types: of its first
  half, partners appended) -- see `_draws_of_path_range`."""
  if skip is None:
    skip = 0
  dtype = np.dtype(np.float32 if dtype is None else dtype)
  batch_shape = tuple(batch_shape or ())
  random_type = RandomType(random_type.value)
  if path_range is not None:
    if batch_shape:
      raise NotImplementedError('path_range with a batch')
    return _draws_of_path_range(num_normal_draws, num_time_steps, num_sample_paths,
                                random_type, skip, seed, dtype, path_range)
  total_dimension = np.zeros(num_time_steps * num_normal_draws, dtype=dtype)
  if random_type in (RandomType.PSEUDO_ANTITHETIC,
                     RandomType.STATELESS_ANTITHETIC):
    sample_shape = (num_sample_paths,) + batch_shape
    is_antithetic = True
  else:
    sample_shape = batch_shape + (num_sample_paths,)
    is_antithetic = False
  draws = mv_normal_sample(sample_shape, mean=total_dimension,
                           random_type=random_type, seed=seed, skip=skip,
                           dtype=dtype)
  draws = draws.reshape(sample_shape + (num_time_steps, num_normal_draws))
  rank = draws.ndim
  if is_antithetic and rank > 3:
    perm = [rank - 2] + list(range(1, rank - 2)) + [0, rank - 1]
  else:
    perm = [rank - 2] + list(range(rank - 2)) + [rank - 1]
  return np.transpose(draws, perm)
