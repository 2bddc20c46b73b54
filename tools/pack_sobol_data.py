"""Packs the Joe-Kuo direction-number file into a compact .npz.

Source: third_party/sobol_data/new-joe-kuo-6.21201 of the reference checkout
(S. Joe and F. Y. Kuo, "Constructing Sobol sequences with better
two-dimensional projections", SIAM J. Sci. Comput. 30, 2635-2654 (2008);
BSD-style licence reproduced in tf-quant-finance_b200/data/SOBOL_LICENSE).
It is DATA (21 200 rows: dimension d, degree s, coefficient a, initial m_i),
not reference source code; the packed arrays are what
`math/random_ops/sobol/sobol_impl.py:237-261` (`load_data`) produces.

  python tools/pack_sobol_data.py [path-to-new-joe-kuo-6.21201]
"""
import os
import sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else (
    '/root/reference/third_party/sobol_data/new-joe-kuo-6.21201')
dst = os.path.join(os.path.dirname(__file__), '..', 'tf-quant-finance_b200',
                   'data', 'joe_kuo_6_21201.npz')
s_arr = np.zeros(21200, dtype=np.uint8)
a_arr = np.zeros(21200, dtype=np.uint32)
m_arr = np.zeros((21200, 18), dtype=np.uint32)
with open(src) as f:
  next(f)
  for k, line in enumerate(f):
    tok = line.split()
    if not tok:
      continue
    s_arr[k] = int(tok[1])
    a_arr[k] = int(tok[2])
    for i, m in enumerate(tok[3:]):
      m_arr[k, i] = int(m)
assert k == 21199, k
np.savez_compressed(dst, s=s_arr, a=a_arr, m=m_arr)
print('wrote', os.path.abspath(dst), os.path.getsize(dst), 'bytes')
