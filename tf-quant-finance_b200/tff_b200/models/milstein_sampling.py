"""Mirror of `tf_quant_finance.models.milstein_sampling` (`milstein_sampling.py:35-256`)
on the B200 path engine: one-dimensional processes with affine coefficients, and
multi-dimensional processes whose volatility does not depend on the state.

Same skeleton as the Euler sampler (`utils.prepare_grid`, coefficients at
`times[i + 1]`, the `_while_loop` recording rule); the step is `_milstein_1d`
(565-575).  The reference precomputes `dim + 3 * dim * stratonovich_order`
normals per step even in one dimension (282-290) and uses the first `dim` of
them: the draw layout is reproduced by generating that tensor on the device and
feeding its first column to the kernel as `normal_draws`.

Multi-dimensional scheme (`_milstein_nd`, 578-595): the higher-order term and the
Stratonovich drift correction are contractions with the volatility GRADIENT
(530-562).  For every multi-dimensional process the engine can run besides the
multi-asset GBM -- drift affine in the state, volatility `B(t)` (`AffineModelND`) --
that gradient is identically zero, the auxiliary `3 * dim * stratonovich_order`
normals multiply zero and the step is `x + dt a + B dW`: the Euler kernel, fed
with the first `dim` columns of the reference's Milstein draw tensor.  Processes
with a state-dependent volatility matrix (SABR, multi-asset GBM) would need the
Stratonovich integrals in the kernel and are refused.
"""
import numpy as np

from tff_b200 import _lib
from tff_b200 import _tensor
from tff_b200 import engine
from tff_b200.math import random
from tff_b200.models import closures
from tff_b200.models import utils


def sample(*, dim, drift_fn, volatility_fn, times, time_step=None, num_time_steps=None,
           num_samples=1, initial_state=None, grad_volatility_fn=None, random_type=None,
           seed=None, swap_memory=True, skip=0, precompute_normal_draws=True,
           watch_params=None, stratonovich_order=5, dtype=None, name=None):
  """Returns sample paths from the process using the Milstein method:
  CUDA tensor `[num_samples, k, dim]`.

  `dim == 1`: the (drift, volatility) pair must be affine in the state (model
  closures or plain Python callables, probed on the host); the volatility
  gradient is then exact and `grad_volatility_fn` is not needed (it is ignored).
  `dim > 1`: drift affine in the state and a state-independent volatility matrix.
  `swap_memory`, `precompute_normal_draws` and `name` only steer TensorFlow's
  execution: the result is defined to equal the reference's precomputed-draws path.
  """
  del swap_memory, precompute_normal_draws, name, grad_volatility_fn
  dim = int(dim)
  del watch_params   # steers TensorFlow's differentiation of the loop only: same forward paths
  dtype = _tensor.infer_dtype(times, dtype)
  times = _tensor.to_numpy(times, dtype).reshape(-1)
  if num_time_steps is not None and time_step is not None:
    raise ValueError('Only one of either `num_time_steps` or `time_step` '
                     'should be defined but not both')
  if time_step is None:
    if num_time_steps is None:
      raise ValueError('Either `num_time_steps` or `time_step` should be defined.')
    num_time_steps = int(num_time_steps)
    time_step = dtype.type(times[-1] / dtype.type(num_time_steps))
  else:
    time_step = dtype.type(_tensor.to_numpy(time_step))
  all_times, keep_mask, _ = utils.prepare_grid(
      times=times, time_step=time_step, num_time_steps=num_time_steps, dtype=dtype)
  if initial_state is None:
    initial_state = np.zeros(dim, dtype=dtype)
  x0 = _tensor.to_numpy(initial_state, dtype).reshape(-1)
  if x0.shape[0] != dim:
    raise NotImplementedError('per-path / batched initial states are not implemented yet')
  euler_spec = closures.resolve_spec(drift_fn, volatility_fn, dim)
  if dim == 1:
    spec = engine.MilsteinSpec1F(euler_spec)
  elif euler_spec.kind == _lib.MODEL_AFFINE_ND:
    spec = euler_spec        # B(t) does not depend on the state: every gradient term is zero
  else:
    raise NotImplementedError(
        'The multi-dimensional B200 Milstein sampler covers processes with affine drift and a '
        'state-independent volatility matrix; a state-dependent volatility needs the Stratonovich '
        'integrals of milstein_sampling.py:481-553 in the kernel. There is no CPU fallback.')
  num_samples = int(num_samples)
  steps_total = all_times.shape[0] - 1
  rt = random.RandomType.PSEUDO if random_type is None else random_type
  draws = utils.generate_mc_normal_draws(
      num_normal_draws=dim + 3 * dim * int(stratonovich_order), num_time_steps=steps_total,
      num_sample_paths=num_samples, random_type=rt, dtype=dtype, seed=seed, skip=skip)
  normal_draws = draws[:, :, :dim].permute(1, 0, 2).contiguous()      # [N, steps, dim]
  del draws
  num_steps, record_slot = engine.record_plan(keep_mask, times.shape[0])
  rng = engine.RngSpec(None, None, 0, normal_draws)
  plan = engine.Plan(spec, all_times, num_steps, x0, rng, num_samples, dtype)
  try:
    return plan.paths(record_slot, times.shape[0])
  finally:
    plan.close()


__all__ = ['sample']
