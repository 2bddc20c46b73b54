"""Oracle (test infrastructure): Brownian-bridge no-touch probabilities.

Restates `black_scholes/brownian_bridge.py`: `brownian_bridge_double` 32-115,
`brownian_bridge_single` 118-196 (numpy, op for op)."""
import numpy as np


def brownian_bridge_double(x_start, x_end, variance, upper_barrier, lower_barrier, n_cutoff=3,
                           dtype=None):
  x_start = np.asarray(x_start, dtype=dtype)
  dtype = x_start.dtype
  variance = np.asarray(variance, dtype=dtype)[..., None]
  x_end = np.asarray(x_end, dtype=dtype)[..., None]
  x_start = x_start[..., None]
  barrier_diff = dtype.type(upper_barrier - lower_barrier)
  k = np.arange(-n_cutoff, n_cutoff + 1, dtype=dtype)[None, :]
  a = k * barrier_diff * (k * barrier_diff + (x_end - x_start))
  b = (k * barrier_diff + x_start - dtype.type(upper_barrier))
  b = b * (k * barrier_diff + (x_end - dtype.type(upper_barrier)))
  return (np.exp(-2 * a / variance) - np.exp(-2 * b / variance)).sum(axis=-1)


def brownian_bridge_single(x_start, x_end, variance, barrier, dtype=None):
  x_start = np.asarray(x_start, dtype=dtype)
  dtype = x_start.dtype
  a = (x_start - dtype.type(barrier)) * (np.asarray(x_end, dtype=dtype) - dtype.type(barrier))
  return 1 - np.exp(-2 * a / np.asarray(variance, dtype=dtype))
