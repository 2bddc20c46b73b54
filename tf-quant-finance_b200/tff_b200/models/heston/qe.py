"""Andersen's Quadratic-Exponential scheme for the Heston model on the device.

Placeholder wired by `HestonModel.sample_paths`; implemented in
`csrc/tqf_paths_kernel.cuh` (HestonQeModel).
"""


def sample_paths(model, times, initial_state, **kwargs):
  raise NotImplementedError('Heston QE scheme: device kernel pending')
