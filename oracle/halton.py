"""Oracle (test infrastructure): the NON-randomized Halton sequence.

Restates `math/random_ops/halton/halton_impl.py:59-288` for `randomized=False`
op for op, in the floating-point arithmetic the reference uses (indices,
weights and digits are tensors of `dtype`; `floor_div`, `%`, `/`, reduce_sum):
  * `_get_indices`            392-413 -> indices = sequence_indices + 1
  * `_MAX_SIZES_BY_AXES`      415-437, 530-534 -> digits kept per axis
  * weights / coeffs / sum    250-288.
The randomized variant (Owen scrambling through TensorFlow's random shuffle) is
not restated: parity unpinned there, and the engine does not implement it.
"""
import numpy as np

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


def sample(dim, num_results=None, sequence_indices=None, dtype=np.float32):
  """`halton.sample(dim, ..., randomized=False)` -> [n, dim] of `dtype`."""
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
  coeffs = coeffs / radixes
  terms = (coeffs / weights).astype(dtype)
  # sequential sum over the coefficient axis in `dtype` (TensorFlow's reduction
  # tree is not specified: parity of the sum is within a few ulp, see tests)
  out = np.zeros(terms.shape[:-1], dtype=dtype)
  for j in range(max_size):
    out = (out + terms[..., j]).astype(dtype)
  return out
