"""Generates tests/golden/reference_tables.npz and tests/golden/oracle_vectors.npz.

  python tests/golden/make_golden.py            (in the builder container: needs /root/reference)

(1) reference_tables.npz -- produced by executing the REFERENCE's own numpy code.
TensorFlow is not installable here, so the two modules below are loaded from
their source files with a stub standing in for `tensorflow` / `tf_quant_finance`
(only their numpy-only functions and module-level tables are used; no TF op runs):
  * math/random_ops/sobol/sobol_impl.py: `load_data()` (237-261) and
    `_compute_direction_numbers(dim)` (171-197) -> direction numbers m[dim][32];
  * math/random_ops/halton/halton_impl.py: `_PRIMES` (440-526) and
    `_MAX_SIZES_BY_AXES` (530-534).
(2) oracle_vectors.npz -- small seeded outputs of oracle/ (itself pinned by the
reference's known-answer tests, tests/test_oracle_kat.py): frozen so that a drift
of the oracle or of the CUDA path shows up against a committed file.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/tf_quant_finance'


class _Stub(types.ModuleType):
  """Any attribute is another stub; stubs are callable and hashable."""

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    child = _Stub(self.__name__ + '.' + name)
    setattr(self, name, child)
    return child

  def __call__(self, *args, **kwargs):
    return _Stub(self.__name__ + '()')


def _load_reference_module(rel_path, name):
  names = ('tensorflow', 'tensorflow.compat', 'tensorflow.compat.v2', 'tf_quant_finance',
           'tf_quant_finance.types', 'tf_quant_finance.math', 'tf_quant_finance.math.random_ops',
           'tf_quant_finance.math.random_ops.stateless')
  saved = {k: sys.modules.get(k) for k in names}
  roots = {'tensorflow': _Stub('tensorflow'), 'tf_quant_finance': _Stub('tf_quant_finance')}
  for name in names:
    parts = name.split('.')
    mod = roots[parts[0]]
    for part in parts[1:]:
      mod = getattr(mod, part)
    mod.__path__ = []            # importable as a package
    sys.modules[name] = mod
  try:
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
  finally:
    for k, v in saved.items():
      if v is None:
        sys.modules.pop(k, None)
      else:
        sys.modules[k] = v


def reference_tables():
  sobol = _load_reference_module('math/random_ops/sobol/sobol_impl.py', '_ref_sobol_impl')
  m = np.asarray(sobol._compute_direction_numbers(256), dtype=np.int32)       # [256, 32]
  # dimensions far into the Joe-Kuo table as well (degree-13+ polynomials)
  m_far = np.asarray(sobol._compute_direction_numbers(16128), dtype=np.int32)[[1000, 5000, 16127]]
  halton = _load_reference_module('math/random_ops/halton/halton_impl.py', '_ref_halton_impl')
  return dict(
      sobol_direction_numbers_256=m,
      sobol_direction_numbers_rows_1000_5000_16127=m_far,
      halton_primes=np.asarray(halton._PRIMES, dtype=np.int32),
      halton_max_sizes_f32=np.asarray(halton._MAX_SIZES_BY_AXES[np.float32]).reshape(-1),
      halton_max_sizes_f64=np.asarray(halton._MAX_SIZES_BY_AXES[np.float64]).reshape(-1))


def oracle_vectors():
  sys.path.insert(0, ROOT)
  from oracle import draws as odraws
  from oracle import euler as oeuler
  from oracle import halton as ohalton
  from oracle import models as omodels
  from oracle import philox as ophilox
  from oracle import sobol as osobol
  out = {}
  out['sobol_points_dim5_skip1000'] = osobol.sample(5, 16, skip=1000, dtype=np.float64)
  out['halton_dim5_idx1000'] = ohalton.sample(5, sequence_indices=np.arange(1000, 1016), dtype=np.float64)
  key, ctr = ophilox.stateless_key_counter([4, 2])
  out['philox_key_counter_seed_4_2'] = np.concatenate([key, ctr]).astype(np.uint32)
  out['philox_raw_words_seed_4_2'] = ophilox.raw_words(key, ctr, 0, 8)
  out['stateless_normal_f64_seed_4_2'] = ophilox.stateless_normal([16], [4, 2], np.float64)
  out['stateless_normal_f32_seed_4_2'] = ophilox.stateless_normal([16], [4, 2], np.float32)
  out['stateless_uniform_f64_seed_4_2'] = ophilox.stateless_uniform([16], [4, 2], np.float64)
  d, v = omodels.heston_closures(2.0, 0.04, 0.5, -0.7, np.float64)
  x0 = np.array([np.log(100.0), 0.04])
  out['heston_euler_sobol_paths'] = oeuler.sample(
      2, d, v, [0.5, 1.0], num_samples=8, initial_state=x0, num_time_steps=4,
      random_type=odraws.RandomType.SOBOL, dtype=np.float64)
  out['heston_euler_stateless_antithetic_paths'] = oeuler.sample(
      2, d, v, [0.5, 1.0], num_samples=8, initial_state=x0, num_time_steps=4,
      random_type=odraws.RandomType.STATELESS_ANTITHETIC, seed=[4, 2], dtype=np.float64)
  return out


if __name__ == '__main__':
  np.savez_compressed(os.path.join(HERE, 'reference_tables.npz'), **reference_tables())
  np.savez_compressed(os.path.join(HERE, 'oracle_vectors.npz'), **oracle_vectors())
  for f in ('reference_tables.npz', 'oracle_vectors.npz'):
    print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')
