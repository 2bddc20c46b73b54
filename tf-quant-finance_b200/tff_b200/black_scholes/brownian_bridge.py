"""Methods for Brownian bridges (`tf_quant_finance/black_scholes/brownian_bridge.py`).

Elementwise no-touch probabilities of a 1-d Brownian bridge, used to move a
discretely monitored barrier payoff to continuous monitoring.  The pricing hot
path evaluates `brownian_bridge_single` step by step inside the fused kernel
(payoffs built with `brownian_bridge=True`, `tff_b200.engine`); the functions
here are the stand-alone API of the reference, evaluated with the framework's
elementwise tensor ops on whatever device the inputs live on.
"""
import numpy as np
import torch

from tff_b200 import _tensor


def _tensors(dtype, *values):
  dt = _tensor.infer_dtype(values[0], dtype, default=np.float32)
  td = _tensor.torch_dtype(dt)
  dev = None
  for v in values:
    if isinstance(v, torch.Tensor):
      dev = v.device
      break
  out = []
  for v in values:
    t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(_tensor.to_numpy(v, dt)))
    out.append(t.to(dtype=td, device=dev) if dev is not None else t.to(dtype=td))
  return out


def brownian_bridge_double(*, x_start, x_end, variance, upper_barrier, lower_barrier,
                           n_cutoff=3, dtype=None, name=None):
  """Probability of not touching either barrier (`brownian_bridge.py:32-115`):
  sum_{k=-n..n} exp(-2 a_k / var) - exp(-2 b_k / var) with
  a_k = k D (k D + x_end - x_start), b_k = (k D + x_start - U)(k D + x_end - U),
  D = U - L."""
  del name
  xs, xe, var = _tensors(dtype, x_start, x_end, variance)
  up, lo = float(upper_barrier), float(lower_barrier)
  diff = up - lo
  xs, xe, var = xs.unsqueeze(-1), xe.unsqueeze(-1), var.unsqueeze(-1)
  k = torch.arange(-int(n_cutoff), int(n_cutoff) + 1, dtype=xs.dtype, device=xs.device).unsqueeze(0)
  a = k * diff * (k * diff + (xe - xs))
  b = (k * diff + xs - up) * (k * diff + (xe - up))
  return (torch.exp(-2 * a / var) - torch.exp(-2 * b / var)).sum(dim=-1)


def brownian_bridge_single(*, x_start, x_end, variance, barrier, dtype=None, name=None):
  """Probability of not touching the barrier (`brownian_bridge.py:118-196`):
  1 - exp(-2 (x_start - B)(x_end - B) / variance) for both ends on one side."""
  del name
  xs, xe, var = _tensors(dtype, x_start, x_end, variance)
  b = float(barrier)
  return 1 - torch.exp(-2 * ((xs - b) * (xe - b)) / var)
