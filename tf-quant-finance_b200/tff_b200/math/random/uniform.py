"""`tff.math.random.uniform` on the device (`math/random_ops/uniform.py:25-153`)."""
import numpy as np

from tff_b200 import _tensor
from tff_b200.math.random import philox
from tff_b200.math.random import sobol
from tff_b200.math.random.multivariate_normal import RandomType


def uniform(dim, sample_shape, random_type=None, dtype=None, seed=None, name=None, **kwargs):
  """Draws from the uniform distribution on [0, 1): CUDA tensor `sample_shape + [dim]`.

  Same contract as the reference: PSEUDO -> `tf.random.uniform(seed=)` (first
  invocation of a fresh op), STATELESS -> `tf.random.stateless_uniform(seed=[a, b],
  alg='philox')`, SOBOL -> `sobol.sample(dim, prod(sample_shape), skip)`;
  PSEUDO_ANTITHETIC raises as in the reference (`uniform.py:102-105`); every other
  type takes the reference's quasi-random branch (`uniform.py:116-153`): HALTON,
  HALTON_RANDOMIZED (`seed` / `randomization_params`) and -- as in the reference,
  whose `else` catches it -- STATELESS_ANTITHETIC, which yields the plain Halton sequence.
  """
  del name
  random_type = RandomType.PSEUDO if random_type is None else random_type
  dtype = _tensor.np_dtype(np.float32 if dtype is None else dtype)
  sample_shape = [int(s) for s in np.asarray(_tensor.to_numpy(sample_shape)).reshape(-1)]
  shape = sample_shape + [int(dim)]
  if random_type.value == RandomType.PSEUDO.value:
    return philox.uniform(shape, dtype=dtype, seed=seed)
  if random_type.value == RandomType.STATELESS.value:
    if seed is None:
      raise ValueError('`seed` must be supplied if the `random_type` is STATELESS.')
    return philox.stateless_uniform(shape, seed, dtype=dtype)
  if random_type.value == RandomType.PSEUDO_ANTITHETIC.value:
    raise NotImplementedError(
        'At the moment antithetic sampling is not supported for the uniform '
        'distribution.')
  if random_type.value == RandomType.SOBOL.value:
    num = int(np.prod(sample_shape)) if sample_shape else 1
    seq = sobol.sample(dim=int(dim), num_results=num, skip=int(kwargs.get('skip', 0)), dtype=dtype)
    return seq.reshape(shape)
  if random_type.value in (RandomType.HALTON.value, RandomType.HALTON_RANDOMIZED.value,
                           RandomType.STATELESS_ANTITHETIC.value):
    # uniform.py:135-150 (the reference's `else` branch)
    from tff_b200.math.random import halton  # pylint: disable=g-import-not-at-top
    num = int(np.prod(sample_shape)) if sample_shape else 1
    skip = int(kwargs.get('skip', 0))
    seq, _ = halton.sample(dim=int(dim), sequence_indices=np.arange(skip, skip + num),
                           randomized=random_type.value == RandomType.HALTON_RANDOMIZED.value,
                           randomization_params=kwargs.get('randomization_params'), seed=seed,
                           dtype=dtype)
    return seq.reshape(shape)
  raise NotImplementedError(
      'uniform: {} is not implemented by the B200 engine.'.format(random_type))
