"""Oracle (test infrastructure): Sobol points in natural order.

Restates `math/random_ops/sobol/sobol_impl.py`:
  * `load_data` 237-261          -> `load_joe_kuo`
  * `_compute_direction_numbers` 171-197 -> `direction_numbers`
  * `sample` 39-167              -> `sample`, `sample_integers`
The direction-number DATA (Joe & Kuo, new-joe-kuo-6.21201) is read from the
packed copy `tf-quant-finance_b200/data/joe_kuo_6_21201.npz` produced by
`tools/pack_sobol_data.py`; `tests/test_oracle_kat.py` checks the packed copy
against the reference's text file whenever /root/reference is present.
"""
import functools
import os
import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..',
                     'tf-quant-finance_b200', 'data', 'joe_kuo_6_21201.npz')


@functools.lru_cache(maxsize=None)
def load_joe_kuo(path=_DATA):
  """Returns (polynomial_coefficients[21200] int64, initial_m[18, 21200])."""
  with np.load(path) as z:
    s = z['s'].astype(np.int64)
    a = z['a'].astype(np.int64)
    m = z['m'].astype(np.int64)
  poly = 2**s + 2 * a + 1                      # sobol_impl.py:257
  return poly, np.ascontiguousarray(m.T)       # [18, 21200] like :251


def parse_joe_kuo_text(path):
  """`load_data` on the original text file (used to check the packed copy)."""
  poly = np.zeros(21200, dtype=np.int64)
  m = np.zeros((18, 21200), dtype=np.int64)
  with open(path) as f:
    next(f)
    index = 0
    for line in f:
      tok = line.split()
      if not tok:
        continue
      poly[index] = 2**int(tok[1]) + 2 * int(tok[2]) + 1
      for i, mi in enumerate(tok[3:]):
        m[i, index] = int(mi)
      index += 1
  return poly, m


@functools.lru_cache(maxsize=8)
def direction_numbers(dim):
  """int64 [dim, 32] matrix of the integers m_{k,j} (`sobol_impl.py:171-197`).

  The reference stores int32; columns j <= 30 (the only ones `sample` can use,
  num_digits <= 31) never overflow, column 31 wraps there and is unused.
  """
  poly, init = load_joe_kuo()
  m = np.zeros((dim, 32), dtype=np.int64)
  m[0, :] = 1
  for k in range(dim - 1):
    a_k = int(poly[k])
    deg = a_k.bit_length() - 1                 # floor(log2(a_k))
    m[k + 1, :deg] = init[:deg, k]
    for j in range(deg, 32):
      v = int(m[k + 1, j - deg])
      for i in range(deg):
        if (a_k >> i) & 1:
          v ^= int(m[k + 1, j - deg + i]) << (deg - i)
      m[k + 1, j] = v & 0xFFFFFFFFFFFF
  return m


def num_digits_for(skip, num_results):
  """`sobol_impl.py:118-123`: ceil(log(max_index) / ln 2) in float64."""
  max_index = int(skip) + int(num_results) + 1
  return int(np.ceil(np.log(np.float64(max_index)) / np.log(2.)))


def sample_integers(dim, num_results, skip=0):
  """int64 [num_results, dim] integer points and num_digits (`:128-157`)."""
  nd = num_digits_for(skip, num_results)
  m = direction_numbers(dim)[:, :nd]
  shifted = m << np.arange(nd - 1, -1, -1, dtype=np.int64)       # [dim, nd]
  irange = np.int64(skip) + 1 + np.arange(num_results, dtype=np.int64)
  out = np.zeros((num_results, dim), dtype=np.int64)
  for b in range(nd):
    bit = (irange >> b) & 1
    out ^= bit[:, None] * shifted[None, :, b]
  return out, nd


def sample(dim, num_results, skip=0, dtype=np.float32):
  """`sobol.sample` (`sobol_impl.py:39-167`): [num_results, dim] in (0, 1)."""
  x, nd = sample_integers(dim, num_results, skip)
  dtype = np.dtype(dtype)
  # int32 -> dtype cast (round to nearest even for float32), then divide by
  # the power of two cast to dtype (`:159-167`).
  return x.astype(dtype) / dtype.type(2**nd)
