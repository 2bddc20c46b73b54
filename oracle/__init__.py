"""CPU oracle for the Euler Monte-Carlo hot path of tf-quant-finance.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`tf-quant-finance_b200/`) may import this package.  The only permitted users
are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py`, and there only as the checker or as the timed
CPU baseline -- never as the thing shipped.

Every function restates, op for op, the algorithm of the reference file cited
in its docstring (paths relative to /root/reference/tf_quant_finance).  The
restatement is numpy / scipy only: TensorFlow is not installable in this
image, so the reference itself cannot be executed.

Pinning status
--------------
* Sobol points, the Sobol->normal draw layout, the time grids, the Heston
  closures, Hull-White bond prices / swaption prices and the Longstaff-Schwartz
  prices are pinned by the reference's own known-answer tests
  (tests/test_oracle_kat.py lists each with its file:line).
* The Philox4x32-10 core is pinned by the Random123 known-answer vectors.
* **parity unpinned**: the TensorFlow-specific part of the pseudo-random
  stream (seed -> key/counter scrambling of `tf.random.stateless_normal`, the
  op-seed pair of `tf.random.normal`, uint32 -> float conversion and the
  Box-Muller constants) lives in TensorFlow's C++ sources, which are neither
  under /root/reference nor installed here, and no reference test holds an
  output value of that stream.  It is restated from the published TensorFlow
  algorithm (tensorflow==2.12.0rc1 is the version pinned by the reference's
  ci_build/Dockerfile:17) -- see oracle/philox.py.
"""
