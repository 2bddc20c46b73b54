"""Oracle (test infrastructure): TensorFlow-compatible Philox4x32-10 normals.

The reference leaves the repository here: `tf.random.stateless_normal(...,
alg='philox')` (`math/random_ops/multivariate_normal.py:268-269`) and
`tf.random.normal(..., seed=)` (`:261-263`) are TensorFlow kernels.  This file
restates the published TensorFlow algorithm (tensorflow==2.12.0rc1, the
version pinned in the reference's `ci_build/Dockerfile:17`):

* `tensorflow/core/lib/random/philox_random.h`       -- Philox4x32-10 core,
  128-bit counter `Skip`, constructor `PhiloxRandom(seed_lo, seed_hi)`.
* `tensorflow/core/kernels/stateless_random_ops.cc`  -- `GenerateKey` (seed
  scrambling with the fixed key 0x3ec8f720 / 0x02461e29).
* `tensorflow/core/lib/random/random_distributions.h` -- `Uint64ToDouble`,
  `Uint32ToFloat`, `BoxMullerDouble`, `BoxMullerFloat`,
  `NormalDistribution<.., double>` (2 outputs per Philox call) and
  `NormalDistribution<.., float>` (4 outputs per call).
* `tensorflow/core/kernels/random_op_cpu.h` -- `FillPhiloxRandomTask`: output
  group g (kResultElementCount elements) is produced from counter + g.
* `tensorflow/python/framework/random_seed.py` -- `get_seed`: with no global
  seed an op seed s becomes the pair (87654321, s).

Pinned: the Philox core by the Random123 known-answer vectors; the float32
stream end to end (GenerateKey, counter layout, Uint32ToFloat, BoxMullerFloat)
by the two vectors TensorFlow prints in its documentation
(tests/test_oracle_kat.py::test_philox_tensorflow_published_*).
PARITY UNPINNED: `Uint64ToDouble` / `BoxMullerDouble` (float64) and the
stateful op-seed pair -- no published float64 value exists and TensorFlow
cannot be run in this image; they are held to internal consistency with the
pinned words only.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

DEFAULT_GRAPH_SEED = 87654321          # random_seed.py


def philox4x32_10(counter, key):
  """Philox4x32-10.  counter: uint32 [..., 4]; key: uint32 [2] -> uint32 [..., 4]."""
  c = np.asarray(counter, dtype=np.uint32).astype(np.uint64)
  c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
  k0, k1 = int(key[0]), int(key[1])
  for _ in range(10):
    p0 = M0 * c0
    p1 = M1 * c2
    n0 = (p1 >> _S32) ^ c1 ^ np.uint64(k0)
    n1 = p1 & _MASK
    n2 = (p0 >> _S32) ^ c3 ^ np.uint64(k1)
    n3 = p0 & _MASK
    c0, c1, c2, c3 = n0, n1, n2, n3
    k0 = (k0 + W0) & 0xFFFFFFFF
    k1 = (k1 + W1) & 0xFFFFFFFF
  return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def counter_add(counter, groups):
  """128-bit `counter + g` for an int array of group numbers (PhiloxRandom::Skip)."""
  g = np.asarray(groups, dtype=np.uint64)
  c = [np.uint64(int(x)) for x in counter]
  lo = c[0] + (g & _MASK)
  r0 = lo & _MASK
  mid = c[1] + (g >> _S32) + (lo >> _S32)
  r1 = mid & _MASK
  hi = c[2] + (mid >> _S32)
  r2 = hi & _MASK
  r3 = (c[3] + (hi >> _S32)) & _MASK
  return np.stack([r0, r1, r2, r3 + np.zeros_like(r0)], axis=-1).astype(np.uint32)


def stateless_key_counter(seed):
  """`GenerateKey`: int seed [2] -> (key uint32[2], counter uint32[4])."""
  s0 = int(seed[0]) & 0xFFFFFFFFFFFFFFFF      # int32/int64 -> uint64 (sign-extends)
  s1 = int(seed[1]) & 0xFFFFFFFFFFFFFFFF
  ctr = np.array([s0 & 0xFFFFFFFF, s0 >> 32, s1 & 0xFFFFFFFF, s1 >> 32],
                 dtype=np.uint32)
  mix = philox4x32_10(ctr, np.array([0x3ec8f720, 0x02461e29], dtype=np.uint32))
  key = np.array([mix[0], mix[1]], dtype=np.uint32)
  counter = np.array([0, 0, mix[2], mix[3]], dtype=np.uint32)
  return key, counter


def stateful_key_counter(op_seed, graph_seed=DEFAULT_GRAPH_SEED):
  """`PhiloxRandom(seed, seed2)` of a FRESH `RandomStandardNormal` kernel.

  `tf.random.normal(seed=s)` with no global seed -> (seed, seed2) =
  (87654321, s) (`random_seed.get_seed`); only the first invocation of a
  fresh kernel is reproducible (each invocation reserves counter space).
  """
  a = int(graph_seed) % (2**31 - 1)          # random_seed._truncate_seed
  b = int(op_seed) % (2**31 - 1)
  if (a, b) == (0, 0):
    a, b = 0, 2**31 - 1
  key = np.array([a & 0xFFFFFFFF, a >> 32], dtype=np.uint32)
  counter = np.array([0, 0, b & 0xFFFFFFFF, b >> 32], dtype=np.uint32)
  return key, counter


def raw_words(key, counter, first_group, num_groups):
  """uint32 [num_groups, 4]: Philox output of groups first_group ..."""
  g = np.uint64(first_group) + np.arange(num_groups, dtype=np.uint64)
  return philox4x32_10(counter_add(counter, g), key)


def uint64_to_double(x0, x1):
  man = ((x0.astype(np.uint64) & np.uint64(0xFFFFF)) << _S32) | x1.astype(np.uint64)
  val = (np.uint64(1023) << np.uint64(52)) | man
  return val.view(np.float64) - 1.0


def uint32_to_float(x):
  val = (np.uint32(127) << np.uint32(23)) | (x.astype(np.uint32) & np.uint32(0x7FFFFF))
  return val.view(np.float32) - np.float32(1.0)


def normals_from_words(words, dtype):
  """`NormalDistribution::operator()` on uint32 [G, 4] -> [G * k] normals."""
  dtype = np.dtype(dtype)
  w = np.ascontiguousarray(words, dtype=np.uint32)
  if dtype == np.float64:
    u1 = np.maximum(uint64_to_double(w[:, 0], w[:, 1]), 1.0e-7)
    v1 = (2 * np.pi) * uint64_to_double(w[:, 2], w[:, 3])
    u2 = np.sqrt(-2.0 * np.log(u1))
    out = np.stack([np.sin(v1) * u2, np.cos(v1) * u2], axis=-1)
    return out.reshape(-1)
  if dtype == np.float32:
    eps = np.float32(1.0e-7)
    two_pi = 2.0 * np.pi      # `2.0f * M_PI * u`: the product is a double
    outs = []
    for i in (0, 2):
      u1 = np.maximum(uint32_to_float(w[:, i]), eps)
      v1 = (two_pi * uint32_to_float(w[:, i + 1]).astype(np.float64)
            ).astype(np.float32)
      u2 = np.sqrt(np.float32(-2.0) * np.log(u1))
      outs += [np.sin(v1) * u2, np.cos(v1) * u2]
    return np.stack(outs, axis=-1).astype(np.float32).reshape(-1)
  raise ValueError(dtype)


def uniforms_from_words(words, dtype):
  """`UniformDistribution::operator()` (random_distributions.h) on uint32 [G, 4]
  -> [G * k] uniforms on [0, 1): float32 four Uint32ToFloat per group, float64
  two Uint64ToDouble.  This is what `tf.random.stateless_uniform` /
  `tf.random.uniform` return for minval 0, maxval 1
  (`math/random_ops/uniform.py:92-101`)."""
  dtype = np.dtype(dtype)
  w = np.ascontiguousarray(words, dtype=np.uint32)
  if dtype == np.float64:
    return np.stack([uint64_to_double(w[:, 0], w[:, 1]), uint64_to_double(w[:, 2], w[:, 3])],
                    axis=-1).reshape(-1)
  if dtype == np.float32:
    return uint32_to_float(w).reshape(-1)
  raise ValueError(dtype)


def uniform_fill(key, counter, num_elements, dtype, first_element=0):
  k = 2 if np.dtype(dtype) == np.float64 else 4
  g0 = first_element // k
  g1 = (first_element + num_elements + k - 1) // k
  flat = uniforms_from_words(raw_words(key, counter, g0, g1 - g0), dtype)
  off = first_element - g0 * k
  return flat[off:off + num_elements]


def stateless_uniform(shape, seed, dtype=np.float32):
  """`tf.random.stateless_uniform(shape, seed, dtype=dtype, alg='philox')`."""
  key, counter = stateless_key_counter(seed)
  return uniform_fill(key, counter, int(np.prod(shape)), dtype).reshape(shape)


def stateful_uniform(shape, seed, dtype=np.float32):
  """First call of `tf.random.uniform(shape, dtype=dtype, seed=seed)`."""
  key, counter = stateful_key_counter(seed)
  return uniform_fill(key, counter, int(np.prod(shape)), dtype).reshape(shape)


def normal_fill(key, counter, num_elements, dtype, first_element=0):
  """Elements [first_element, first_element + num_elements) of the flat stream."""
  k = 2 if np.dtype(dtype) == np.float64 else 4
  g0 = first_element // k
  g1 = (first_element + num_elements + k - 1) // k
  flat = normals_from_words(raw_words(key, counter, g0, g1 - g0), dtype)
  off = first_element - g0 * k
  return flat[off:off + num_elements]


def stateless_normal(shape, seed, dtype=np.float32):
  """`tf.random.stateless_normal(shape, seed, dtype, alg='philox')`."""
  key, counter = stateless_key_counter(seed)
  n = int(np.prod(shape))
  return normal_fill(key, counter, n, dtype).reshape(shape)


def stateful_normal(shape, seed, dtype=np.float32):
  """First call of `tf.random.normal(shape, dtype=dtype, seed=seed)`."""
  key, counter = stateful_key_counter(seed)
  n = int(np.prod(shape))
  return normal_fill(key, counter, n, dtype).reshape(shape)
