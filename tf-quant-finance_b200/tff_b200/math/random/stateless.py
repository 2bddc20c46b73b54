"""`stateless_random_shuffle` on the device (`math/random_ops/stateless.py:24-52`)."""
import numpy as np
import torch

from tff_b200 import _tensor
from tff_b200.math.random import philox


def stateless_random_shuffle(input_tensor, seed, name=None):
  """Stateless random shuffle of the first dimension of `input_tensor`.

  As the reference: one `tf.random.stateless_uniform([n], seed, float64)` draw
  per row (the engine's bit-exact Philox uniforms) and a STABLE argsort of them;
  run twice with the same seed it returns the same permutation, independent of
  the values and dtype of the input.  Returns a CUDA tensor of the input's
  shape and dtype.
  """
  del name
  if isinstance(input_tensor, torch.Tensor) or hasattr(input_tensor, '__dlpack__'):
    x = _tensor.from_dlpack(input_tensor).to(_tensor.device())
  else:
    x = torch.as_tensor(np.asarray(input_tensor), device=_tensor.device())
  n = int(x.shape[0])
  uniforms = philox.stateless_uniform([n], seed, dtype=np.float64)
  order = torch.argsort(uniforms, stable=True)
  return x.index_select(0, order)
