"""Compare what the HOST side hands to the device between a past commit and the working tree.

  python tools/dev/compare_host_tables.py <commit>

Without a GPU a change to the host code (grids, piecewise parameters, coefficient tables) cannot be
re-validated against the kernels; but if every table the host builds is BIT-IDENTICAL to the one built
by a commit whose GPU suite was green, the device sees the same inputs.  The script extracts
`tf-quant-finance_b200/tff_b200` and `oracle/` of <commit> into a temporary directory, builds a fixed
set of tables / oracle outputs with both trees (piecewise-constant parameters on a random grid, both
dtypes) and prints which are equal bit for bit.  CPU only; the built `libtqf.so` of the working tree is
symlinked into the old tree (it is only loaded for its constants).
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _load(root):
  for m in [k for k in sys.modules if k.split('.')[0] in ('tff_b200', 'oracle')]:
    del sys.modules[m]
  sys.path[:0] = [root + '/tf-quant-finance_b200', root]
  from oracle import draws as od
  from oracle import models as om
  from tff_b200 import engine
  from tff_b200.math import piecewise
  from tff_b200.models.geometric_brownian_motion import exact
  from tff_b200.models.heston import qe
  from tff_b200.models.hull_white import one_factor
  del sys.path[:2]
  return od, om, engine, piecewise, exact, qe, one_factor


def _outputs(root):
  od, om, engine, piecewise, exact, qe, one_factor = _load(root)
  out = {}
  rs = np.random.RandomState(0)
  for dtype in (np.float64, np.float32):
    n = np.dtype(dtype).name
    pw = lambda j, v: piecewise.PiecewiseConstantFunc(j, v, dtype=dtype)
    t = np.sort(rs.uniform(0, 2, 40)).astype(dtype)
    out['heston_euler_table/' + n] = engine.HestonEulerSpec(pw([0.5], [1.0, 1.1]), 0.04, pw([0.3], [0.5, 0.8]),
                                                            -0.7).coef_table(t, dtype)
    out['heston_qe_table/' + n] = qe.HestonQeSpec(pw([0.5], [1.0, 1.1]), 0.04, pw([0.3], [0.5, 0.8]), -0.7,
                                                  1e-6).coef_table(t, dtype)
    out['gbm_table/' + n] = engine.GbmSpec1F(pw([0.3], [0.05, 0.02]), 0.3).coef_table(t, dtype)
    f = pw([0.3, 0.8], [0.1, 0.2, 0.15])
    out['piecewise_value/' + n] = np.asarray(f(t))
    out['piecewise_integral/' + n] = np.asarray(f.integrate(t[:-1], t[1:]), dtype=dtype)
    out['gbm_exact_vol2_integral/' + n] = np.asarray(exact._integrate(f, t[:-1], t[1:], dtype, square=True))
    m = one_factor.HullWhiteModel1F(0.03, pw([0.5, 1.5], [0.01, 0.02, 0.015]),
                                    lambda x: 0.01 * np.ones_like(np.asarray(x)), dtype=dtype)
    at, _, _ = m._prepare_grid(np.array([0.25, 0.5, 1.0, 2.0], dtype), None)
    out['hw1f_table/' + n] = one_factor.HullWhite1FSpec(m._tables, m._fwd, None).coef_table(at, dtype)
    ovol = om.PiecewiseConstantFunc([0.3, 0.8], [0.1, 0.2, 0.15], dtype=dtype)
    out['oracle_gbm_exact/' + n] = om.gbm_exact_sample_paths(0.03, ovol, np.array([0.1, 0.5, 1.0, 2.0], dtype),
                                                             np.array([1.5], dtype), 256, od.RandomType.SOBOL, None, 5,
                                                             dtype)
    out['oracle_mv_normal/' + n] = od.mv_normal_sample([500], np.zeros(6, dtype), random_type=od.RandomType.HALTON,
                                                       skip=3)
    out['oracle_batched_draws/' + n] = od.generate_mc_normal_draws(2, 5, 16, od.RandomType.STATELESS_ANTITHETIC,
                                                                   batch_shape=(3,), seed=[1, 2], dtype=dtype)
  return out


def main():
  commit = sys.argv[1]
  tmp = tempfile.mkdtemp()
  try:
    subprocess.check_call('git archive %s tf-quant-finance_b200/tff_b200 tf-quant-finance_b200/data oracle '
                          '| tar -x -C %s' % (commit, tmp), shell=True, cwd=ROOT)
    lib = tmp + '/tf-quant-finance_b200/tff_b200/lib'
    os.makedirs(lib, exist_ok=True)
    os.symlink(ROOT + '/tf-quant-finance_b200/tff_b200/lib/libtqf.so', lib + '/libtqf.so')
    old, new = _outputs(tmp), _outputs(ROOT)
  finally:
    shutil.rmtree(tmp)
  bad = 0
  for k in sorted(new):
    same = k in old and old[k].dtype == new[k].dtype and old[k].shape == new[k].shape and np.array_equal(
        old[k], new[k], equal_nan=True)
    bad += not same
    print('%-34s %s' % (k, 'identical' if same else 'DIFFERENT'))
  sys.exit(1 if bad else 0)


if __name__ == '__main__':
  main()
