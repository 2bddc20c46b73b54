"""Exact log-normal GBM samplers on the device (pending)."""


def sample_paths_univariate(model, times, **kwargs):
  raise NotImplementedError('exact GBM sampler: device kernel pending')
