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
* The TensorFlow-specific part of the float32 pseudo-random stream (seed ->
  key/counter scrambling of `tf.random.stateless_normal`, the group / counter
  layout, `Uint32ToFloat`, `BoxMullerFloat`) is pinned by the two vectors
  TensorFlow publishes in its own documentation (`tf.random.stateless_normal(
  [2, 3], seed=[1, 2])` and `tf.random.Generator.from_seed(1).normal([2, 3])`),
  tests/test_oracle_kat.py::test_philox_tensorflow_published_*.
* TensorFlow's stateless INTEGER uniform (`lo + x mod range`) is pinned by the
  reference's documented `random_digital_shift(2, 10, seed=(2, 3)) == [586, 1011]`
  (math/qmc/digital_net.py:58-68), tests/test_qmc.py; `tff.math.qmc` as a whole
  by every known value of its three test files.
* **parity unpinned** (only this): the float64 conversion `Uint64ToDouble` /
  `BoxMullerDouble` and the op-seed pair `(87654321, s)` of the stateful
  `tf.random.normal(seed=s)`.  TensorFlow cannot be installed in this image and
  neither its documentation nor any test of the reference prints a float64
  value of the stream.  They are restated from the published TensorFlow source
  (tensorflow==2.12.0rc1, the version pinned by the reference's
  ci_build/Dockerfile:17) in oracle/philox.py and held to what CAN be checked
  without TensorFlow: the float64 stream consumes the same (pinned) raw words,
  its uniforms have the documented bit layout and its normals invert back to
  those uniforms (z0^2 + z1^2 = -2 ln u1, atan2(z0, z1) = 2 pi u2).
"""
