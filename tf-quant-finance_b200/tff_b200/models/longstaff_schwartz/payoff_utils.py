"""Payoff descriptors for the LSM passes
(`tf_quant_finance/models/longstaff_schwartz/payoff_utils.py:27-97`)."""
import numpy as np
import torch

from tff_b200 import _tensor


class BasketPutPayoff:
  """relu(strikes - mean_dim x) -> `[num_samples, batch_size]`.

  Callable on the host with torch tensors like the reference's `put_valuer`;
  `least_square_mc` recognises it and evaluates it inside the fused passes."""

  def __init__(self, strikes, dtype=None):
    dt = _tensor.infer_dtype(strikes, dtype, default=np.float32)
    self.strikes = np.atleast_1d(_tensor.to_numpy(strikes, dt))
    if self.strikes.ndim != 1:
      raise NotImplementedError(
          'per-sample strikes (`[num_samples, batch_size]`) are not '
          'implemented by the B200 engine; pass `[batch_size]` strikes.')
    self.dtype = dt

  def __call__(self, sample_paths, time_index):
    x = sample_paths if isinstance(sample_paths, torch.Tensor) else torch.as_tensor(
        np.asarray(sample_paths))
    x = x.unsqueeze(1) if x.dim() == 3 else x.permute(1, 0, 2, 3)
    sl = x[:, :, int(time_index), :]
    k = torch.as_tensor(self.strikes, dtype=x.dtype, device=x.device)
    return torch.relu(k - sl.mean(dim=-1))


def make_basket_put_payoff(strikes, dtype=None, name=None):
  """Produces the payoff of a simple basket put option (`payoff_utils.py:27-58`)."""
  del name
  return BasketPutPayoff(strikes, dtype)


class TabulatedPayoff:
  """Exercise values known in advance: `values[time_index]` is the
  `[num_samples, batch_size]` payoff of exercising at `time_index`.

  The reference's Bermudan swaption pricer hands `least_square_mc` a closure
  over such a precomputed tensor (`hull_white/swaption.py:698-706`); a Python
  closure cannot run inside the fused LSM passes, this descriptor can: the
  passes read the values from device memory."""

  def __init__(self, values):
    self.values = _tensor.from_dlpack(values) if not isinstance(values, torch.Tensor) else values
    if self.values.dim() != 3:
      raise ValueError('values must have shape [num_times, num_samples, batch_size]')

  def __call__(self, sample_paths, time_index):
    del sample_paths
    return self.values[int(time_index)]


def make_tabulated_payoff(values):
  """Payoff descriptor of precomputed exercise values `[num_times, N, B]`."""
  return TabulatedPayoff(values)
