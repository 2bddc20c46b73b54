"""TEST INFRASTRUCTURE (CPU oracle) - numpy restatement of `tff.math.qmc`.

Follows, op for op, the reference's
  math/qmc/utils.py:23-158          exp2 / log2 / tent_transform / filter_tensor
  math/qmc/digital_net.py:45-527    random_digital_shift, random_scrambling_matrices,
                                    digital_net_sample, scramble_generating_matrices
  math/qmc/sobol.py:32-395          sobol_sample, sobol_generating_matrices
  math/qmc/lattice_rule.py:40-229   random_scrambling_vectors, lattice_rule_sample
(the `tf.while_loop`s become Python loops over the same loop variables).

Pinned by the reference's own known values: the 29 x 5 Sobol table, the
sequence-index and tent-transform tables, the generating matrices
`[[16, 8, 4, 2, 1], [16, 24, 20, 30, 17], ...]` (sobol_test.py:64-190,
digital_net_test.py:104-245), the lattice tables (lattice_rule_test.py) and the
documented `random_digital_shift(2, 10, seed=(2, 3)) == [586, 1011]`
(digital_net.py:58-68), which also pins TensorFlow's stateless integer
uniform on top of `oracle/philox.py`.

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import
this module; the product path never does.
"""
import numpy as np

from oracle import philox
from oracle import sobol as _sobol


# ------------------------------------------------------------------ utils ----
def exp2(value, dtype=np.int32):
  """utils.py:23-54: `1 << value`, saturated to dtype.max."""
  dtype = np.dtype(dtype)
  value = np.asarray(value, dtype=dtype)
  limit = 8 * dtype.itemsize - (0 if dtype.kind == 'u' else 1)
  safe = np.where(value >= limit, 0, value).astype(dtype)
  return np.where(value >= limit, np.iinfo(dtype).max,
                  np.left_shift(np.ones_like(value), safe)).astype(dtype)


def log2(value):
  """utils.py:57-77: log(value) / log(2) in the dtype of `value`."""
  value = np.asarray(value)
  return np.log(value) / np.log(np.asarray(2, dtype=value.dtype))


def _ceil_log2_f32(n):
  return int(np.ceil(log2(np.float32(n))))


def tent_transform(value):
  """utils.py:94-117."""
  return np.where(value < 0.5, 2 * value, 2 * (1 - value)).astype(value.dtype)


def filter_tensor(value, bit_mask, bit_index):
  """utils.py:120-158: `value` where bit `bit_index` of `bit_mask` is set, else 0."""
  value = np.asarray(value)
  shift = np.clip(np.asarray(bit_index), 0, 8 * value.dtype.itemsize - 1)
  bit = (np.right_shift(np.asarray(bit_mask, dtype=value.dtype), shift.astype(value.dtype)) & 1) == 1
  bit = bit & (np.asarray(bit_index) >= 0)
  return np.where(bit, value, 0).astype(value.dtype)


# -------------------------------------------------- stateless int uniform ----
def stateless_uniform_int(shape, seed, minval, maxval, dtype=np.int32):
  """`tf.random.stateless_uniform(shape, seed, minval, maxval, dtype=int)`.

  tensorflow/core/lib/random/random_distributions.h,
  `UniformDistribution<PhiloxRandom, int32>`: `lo + x % (hi - lo)` on every
  uint32 word (four per Philox call); `int64`: two words per value,
  `lo + (x0 | x1 << 32) % (hi - lo)`.
  """
  dtype = np.dtype(dtype)
  n = int(np.prod(shape))
  key, counter = philox.stateless_key_counter(seed)
  if dtype == np.int32:
    words = philox.raw_words(key, counter, 0, (n + 3) // 4).reshape(-1)[:n].astype(np.uint64)
    rng = np.uint64((int(maxval) - int(minval)) & 0xFFFFFFFF)
    out = (np.int64(minval) + (words % rng).astype(np.int64)).astype(np.int32)
  elif dtype == np.int64:
    words = philox.raw_words(key, counter, 0, (n + 1) // 2).reshape(-1, 2)[:n].astype(np.uint64)
    x = words[:, 0] | (words[:, 1] << np.uint64(32))
    rng = np.uint64((int(maxval) - int(minval)) & 0xFFFFFFFFFFFFFFFF)
    out = (np.uint64(int(minval) & 0xFFFFFFFFFFFFFFFF) + x % rng).astype(np.int64)
  else:
    raise ValueError('dtype must be int32 or int64')
  return out.reshape(shape)


def _random_stateless_uniform(shape, num_digits, seed, dtype):
  """digital_net.py:155-202."""
  dtype = np.dtype(dtype or np.int32)
  minval = exp2(np.asarray(num_digits, dtype=dtype) - 1, dtype)
  maxval = exp2(np.asarray(num_digits, dtype=dtype), dtype)
  return stateless_uniform_int(shape, seed, int(minval), int(maxval), dtype)


def random_digital_shift(dim, num_digits, seed, dtype=None):
  """digital_net.py:45-95."""
  return _random_stateless_uniform((int(dim),), num_digits, seed, dtype)


def random_scrambling_matrices(dim, num_digits, seed, dtype=None):
  """digital_net.py:98-152."""
  return _random_stateless_uniform((int(dim), int(num_digits)), num_digits, seed, dtype)


# ------------------------------------------------------------ digital net ----
def scramble_generating_matrices(generating_matrices, scrambling_matrices, num_digits, dtype=None):
  """digital_net.py:422-527."""
  generating_matrices = np.asarray(generating_matrices)
  dtype = np.dtype(dtype or generating_matrices.dtype)
  g = generating_matrices.astype(dtype)
  s = np.asarray(scrambling_matrices).astype(dtype)
  matrix = np.zeros_like(g)
  for shift in range(int(num_digits)):
    shifted = np.right_shift(s[:, shift:shift + 1], dtype.type(shift))
    matrix = matrix ^ filter_tensor(shifted, g, int(num_digits) - 1 - shift)
  return matrix


def digital_net_sample(generating_matrices, num_results, num_digits, sequence_indices=None,
                       scrambling_matrices=None, digital_shift=None, apply_tent_transform=False,
                       dtype=None):
  """digital_net.py:205-419."""
  g = np.asarray(generating_matrices)
  int_dtype = g.dtype
  real_dtype = np.dtype(dtype or np.float32)
  dim = g.shape[0]
  log_num_results = _ceil_log2_f32(num_results)
  if sequence_indices is None:
    sequence_indices = np.arange(0, int(num_results), dtype=int_dtype)
  idx = np.asarray(sequence_indices).astype(int_dtype)
  if digital_shift is None:
    digital_shift = np.zeros(dim, dtype=int_dtype)
  digital_shift = np.asarray(digital_shift).astype(int_dtype)
  if scrambling_matrices is not None:
    g = scramble_generating_matrices(g, scrambling_matrices, num_digits, dtype=int_dtype)
  points = np.repeat(digital_shift[None, :], idx.size, axis=0)
  for log_index in range(log_num_results):
    points = points ^ filter_tensor(g[None, :, log_index], idx[:, None], log_index)
  max_binary_point = np.left_shift(int_dtype.type(1), int_dtype.type(num_digits))
  out = points.astype(real_dtype) / real_dtype.type(max_binary_point)
  return tent_transform(out) if apply_tent_transform else out


# ------------------------------------------------------------------ sobol ----
def _identity_matrix(num_columns, num_digits, dtype):
  """sobol.py:221-243."""
  shifts = np.arange(num_digits - 1, num_digits - 1 - num_columns, -1)
  return np.left_shift(np.ones((1, num_columns), dtype=dtype), shifts.astype(dtype))


def _sobol_generating_matrices(dim, log_num_results, num_digits, dtype):
  """sobol.py:246-395."""
  poly_all, init_all = _sobol.load_joe_kuo()
  dtype = np.dtype(dtype)
  indices = np.arange(log_num_results).astype(dtype)
  directions = init_all
  pad = max(0, log_num_results - directions.shape[0])
  directions = np.pad(directions, [[0, pad], [0, 0]])[:log_num_results]
  directions = directions[:, :dim].T.astype(dtype)                   # [dim, log_num_results]
  polynomial = poly_all[:dim, None].astype(dtype)
  degree = np.floor(log2(polynomial.astype(np.float32))).astype(dtype)
  matrices = np.left_shift(directions, (num_digits - 1 - indices)[None, :].astype(dtype))
  for column in range(log_num_results - 1):
    column_values = matrices[:, column:column + 1]
    should = (np.maximum(degree, column + 1) <= indices) & (indices <= column + degree)
    updated = np.where(indices == column + degree, np.right_shift(column_values, degree), matrices) \
        ^ filter_tensor(column_values, polynomial, column + degree - indices)
    matrices = np.where(should, updated, matrices).astype(dtype)
  return matrices


def sobol_generating_matrices(dim, num_results, num_digits, dtype=None):
  """sobol.py:132-218."""
  dtype = np.dtype(dtype or np.int32)
  log_num_results = _ceil_log2_f32(num_results)
  identity = _identity_matrix(log_num_results, int(num_digits), dtype)
  if int(dim) == 1:
    return identity
  matrices = _sobol_generating_matrices(int(dim) - 1, log_num_results, int(num_digits), dtype)
  return np.concatenate([identity, matrices], axis=0)


def sobol_sample(dim, num_results, sequence_indices=None, digital_shift=None,
                 scrambling_matrices=None, apply_tent_transform=False, dtype=None):
  """sobol.py:32-129."""
  num_digits = _ceil_log2_f32(num_results)
  g = sobol_generating_matrices(dim, num_results, num_digits, dtype=np.int32)
  if scrambling_matrices is not None:
    g = scramble_generating_matrices(g, scrambling_matrices, num_digits)
  return digital_net_sample(g, num_results, num_digits, sequence_indices=sequence_indices,
                            digital_shift=digital_shift, apply_tent_transform=apply_tent_transform,
                            dtype=dtype or np.float32)


# ----------------------------------------------------------- lattice rule ----
def random_scrambling_vectors(dim, seed, dtype=None):
  """lattice_rule.py:40-96: stateless uniforms in [0, 1)."""
  return philox.stateless_uniform((int(dim),), seed, dtype=np.dtype(dtype or np.float32))


def lattice_rule_sample(generating_vectors, dim, num_results, sequence_indices=None,
                        additive_shift=None, apply_tent_transform=False, dtype=None):
  """lattice_rule.py:99-229."""
  gv = np.asarray(generating_vectors)
  int_dtype = gv.dtype
  real = np.dtype(dtype or np.float32)
  dim = int(dim)
  if sequence_indices is None:
    sequence_indices = np.arange(0, int(num_results))
  idx = np.asarray(sequence_indices).astype(int_dtype)
  unit = real.type(1)
  scaled = gv[:dim].astype(real) / real.type(num_results)
  points = idx.astype(real)[:, None] * np.mod(scaled, unit)[None, :]
  if additive_shift is not None:
    points = points + np.asarray(additive_shift).astype(real)[:dim]
  points = np.mod(points, unit).astype(real)
  return tent_transform(points) if apply_tent_transform else points
