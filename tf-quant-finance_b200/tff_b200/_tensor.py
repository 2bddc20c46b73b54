"""Tensor plumbing: dtypes, device buffers and DLPack interchange.

PyTorch is used for device memory, streams and torch.distributed only.  Any
object exporting `__dlpack__` (a `tf.Tensor` through
`tf.experimental.dlpack`, a cupy array, ...) is accepted zero-copy wherever a
device tensor is expected, and results can be handed back to TensorFlow with
`to_tf` when TensorFlow is installed.
"""
import numpy as np
import torch

from tff_b200 import _lib


def np_dtype(dtype, default=np.float32):
  """numpy dtype from numpy / torch / tf dtypes, `None` -> default."""
  if dtype is None:
    return np.dtype(default)
  if isinstance(dtype, torch.dtype):
    return np.dtype({torch.float32: np.float32, torch.float64: np.float64,
                     torch.int32: np.int32, torch.int64: np.int64}[dtype])
  if hasattr(dtype, 'as_numpy_dtype'):      # tf.DType
    return np.dtype(dtype.as_numpy_dtype)
  return np.dtype(dtype)


def torch_dtype(dtype):
  return {np.dtype(np.float32): torch.float32,
          np.dtype(np.float64): torch.float64,
          np.dtype(np.int32): torch.int32,
          np.dtype(np.int64): torch.int64,
          np.dtype(np.uint32): torch.uint32}[np.dtype(dtype)]


def tqf_dtype(dtype):
  dtype = np.dtype(dtype)
  if dtype == np.float64:
    return _lib.F64
  if dtype == np.float32:
    return _lib.F32
  raise ValueError('dtype must be float32 or float64, got {}'.format(dtype))


def infer_dtype(value, dtype=None, default=np.float32):
  """dtype following `tf.convert_to_tensor(value, dtype)` conventions."""
  if dtype is not None:
    return np_dtype(dtype)
  if isinstance(value, torch.Tensor):
    return np_dtype(value.dtype)
  if isinstance(value, np.ndarray) and value.dtype.kind == 'f':
    return value.dtype
  if isinstance(value, np.floating):
    return np.dtype(type(value))
  return np.dtype(default)


def to_numpy(value, dtype=None):
  """Host numpy copy of a (small) parameter given as anything array-like."""
  if isinstance(value, torch.Tensor):
    value = value.detach().cpu().numpy()
  elif hasattr(value, 'numpy') and not isinstance(value, np.ndarray):
    value = value.numpy()                   # tf.Tensor (eager)
  return np.asarray(value, dtype=dtype)


def device():
  _lib.require_cuda()
  return torch.device('cuda', torch.cuda.current_device())


def current_stream_ptr():
  return torch.cuda.current_stream().cuda_stream


def empty(shape, dtype):
  return torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(dtype),
                     device=device())


def from_dlpack(x):
  """Zero-copy torch view of any DLPack-exporting device tensor."""
  if isinstance(x, torch.Tensor):
    return x
  if hasattr(x, '__dlpack__'):
    return torch.from_dlpack(x)
  raise TypeError('expected a torch.Tensor or a DLPack-exporting tensor')


def to_tf(t):
  """Zero-copy `tf.Tensor` view of a result (requires TensorFlow)."""
  import tensorflow as tf  # pylint: disable=g-import-not-at-top
  return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))
