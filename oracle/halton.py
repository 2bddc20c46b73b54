"""Oracle (test infrastructure): the Halton sequence, plain and randomized.

Restates `math/random_ops/halton/halton_impl.py:59-322` op for op, in the
floating-point arithmetic the reference uses (indices, weights and digits are
tensors of `dtype`; `floor_div`, `%`, `/`, reduce_sum):
  * `_get_indices`            392-413 -> indices = sequence_indices + 1
  * `_MAX_SIZES_BY_AXES`      415-437, 530-534 -> digits kept per axis
  * weights / coeffs / sum    250-288
  * `_randomize`              325-339, `_get_permutations` 342-379 (Owen 2017):
    digit j of axis d goes through its own permutation of range(p_d), drawn by
    `stateless_random_shuffle(range(p_d), seed=(seed + j, p_d))`
    (`math/random_ops/stateless.py:24-52`: float64 `stateless_uniform` + stable
    argsort); the trailing zero digits are accounted for by `zero_correction`
    = `stateless_uniform([dim, 1], seed=(seed, seed), dtype) / p_d^size_d` (303-322).
The non-randomized values are pinned by `halton_test.py:30-61`.  The randomized
variant has no output value in the reference's tests (they are statistical):
beyond its structure (permutations, range, the zero correction, reuse of
`randomization_params`) it rests on the Philox pins of `oracle/philox.py` --
float64 `stateless_uniform` conversion: parity unpinned.
"""
import numpy as np

from oracle import philox

MAX_DIMENSION = 1000
MAX_INDEX_BY_DTYPE = {np.dtype(np.float32): 2**24 - 1, np.dtype(np.float64): 2**53 - 1}


def primes(n):
  """The first `n` primes (`_PRIMES`, halton_impl.py:440-526, is the first 1000)."""
  out, c = [], 2
  while len(out) < n:
    if all(c % p for p in out if p * p <= c):
      out.append(c)
    c += 1
  return np.array(out, dtype=np.int32)


def max_sizes_by_axes(dim, dtype):
  """`_base_expansion_size(_MAX_INDEX_BY_DTYPE[dtype], _PRIMES)` (415-437)."""
  # The reference evaluates this at import time on a Python int and the int32
  # primes, i.e. in float64 whatever `dtype` is (24 digits in base 2 for
  # float32, 54 for float64 -- `_NUM_COEFFS_BY_DTYPE`).
  dtype = np.dtype(dtype)
  num = MAX_INDEX_BY_DTYPE[dtype]
  bases = primes(dim).reshape(dim, 1)
  return (np.floor(np.log(num) / np.log(bases)) + 1).astype(dtype)      # [dim, 1]


def num_coeffs(dtype):
  """`_NUM_COEFFS_BY_DTYPE` (halton_impl.py:46): digits of the largest index in base 2."""
  return {np.dtype(np.float32): 24, np.dtype(np.float64): 54}[np.dtype(dtype)]


def stateless_random_shuffle(values, seed):
  """`stateless.py:24-52`: gather by the stable argsort of float64 stateless uniforms."""
  values = np.asarray(values)
  u = philox.stateless_uniform([values.shape[0]], seed, np.float64)
  return values[np.argsort(u, kind='stable')]


def get_permutations(num_results, dims, seed):
  """`_get_permutations` (342-379) -> int32 [num_results, sum(dims)]."""
  cols = []
  for d in dims:
    d = int(d)
    cols.append(np.stack([stateless_random_shuffle(np.arange(d, dtype=np.int32), (seed + i, d))
                          for i in range(num_results)], 0))
  return np.concatenate(cols, axis=-1)


def zero_correction(dim, seed, dtype):
  """303-322: stateless_uniform([dim, 1], (seed, seed)) / radixes**max_sizes_by_axes."""
  dtype = np.dtype(dtype)
  radixes = primes(dim).astype(dtype).reshape(dim, 1)
  u = philox.stateless_uniform([dim, 1], (seed, seed), dtype)
  return (u / (radixes**max_sizes_by_axes(dim, dtype)).astype(dtype)).astype(dtype).reshape(-1)


def sample(dim, num_results=None, sequence_indices=None, dtype=np.float32, randomized=False,
           seed=None, randomization_params=None, return_params=False):
  """`halton.sample(dim, ...)` -> [n, dim] of `dtype` (and `(perms, zero_correction)` with
  `return_params`).  `randomized=True` needs an integer `seed` or `randomization_params`."""
  if (num_results is None) == (sequence_indices is None):
    raise ValueError('Either `num_results` or `sequence_indices` must be'
                     ' specified but not both.')
  dtype = np.dtype(dtype)
  if sequence_indices is None:
    sequence_indices = np.arange(int(num_results))
  indices = (np.asarray(sequence_indices).astype(dtype) + dtype.type(1)).reshape(-1, 1, 1)
  radixes = primes(dim).astype(dtype).reshape(dim, 1)
  sizes = max_sizes_by_axes(dim, dtype)
  max_size = int(sizes.max())
  exponents = np.tile(np.arange(max_size, dtype=dtype)[None, :], [dim, 1])
  weight_mask = exponents >= sizes
  capped = np.where(weight_mask, np.zeros_like(exponents), exponents)
  weights = np.round(radixes**capped).astype(dtype)
  coeffs = np.floor_divide(indices, weights)
  coeffs = coeffs * (dtype.type(1) - weight_mask.astype(dtype))
  coeffs = np.mod(coeffs, radixes)
  params = None
  if randomized:
    if randomization_params is None:
      perms, zc = None, None
    else:
      perms, zc = randomization_params
    # _randomize (325-339)
    nc = num_coeffs(dtype)
    rad = primes(dim).astype(np.int32)
    if perms is None:
      perms = get_permutations(nc, rad, int(seed)).reshape(-1)
    radix_sum = int(rad.sum())
    radix_offsets = (np.cumsum(rad) - rad).reshape(-1, 1)
    offsets = radix_offsets + np.arange(nc) * radix_sum                 # [dim, nc]
    coeffs = np.asarray(perms)[coeffs.astype(np.int32) + offsets].astype(dtype)
    coeffs = coeffs * (dtype.type(1) - weight_mask.astype(dtype))
    if zc is None:
      zc = zero_correction(dim, int(seed), dtype)
    params = (perms, zc)
  coeffs = coeffs / radixes
  terms = (coeffs / weights).astype(dtype)
  # sequential sum over the coefficient axis in `dtype` (TensorFlow's reduction
  # tree is not specified: parity of the sum is within a few ulp, see tests)
  out = np.zeros(terms.shape[:-1], dtype=dtype)
  for j in range(max_size):
    out = (out + terms[..., j]).astype(dtype)
  if randomized:
    out = (out + np.asarray(params[1], dtype=dtype)).astype(dtype)
  return (out, params) if return_params else out
