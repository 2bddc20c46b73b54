"""Interface of an Ito process (`tf_quant_finance/models/ito_process.py:40-460`).

Only the sampling part of the interface is on the hot path; the PDE hooks
(`fd_solver_backward` / `fd_solver_forward`) are out of scope.
"""
import abc


class ItoProcess(abc.ABC):
  """dX_i = a_i(t, X) dt + Sum_j S_ij(t, X) dW_j."""

  @abc.abstractmethod
  def name(self):
    """The name to give to ops created by this class."""

  @abc.abstractmethod
  def dim(self):
    """The dimension of the process."""

  @abc.abstractmethod
  def dtype(self):
    """The data type of process realizations."""

  @abc.abstractmethod
  def drift_fn(self):
    """Python callable calculating instantaneous drift."""

  @abc.abstractmethod
  def volatility_fn(self):
    """Python callable calculating the instantaneous volatility matrix."""

  @abc.abstractmethod
  def sample_paths(self, times, num_samples=1, initial_state=None,
                   random_type=None, seed=None, **kwargs):
    """Returns a sample of paths from the process."""

  def fd_solver_backward(self, *args, **kwargs):
    raise NotImplementedError('PDE solvers are outside the B200 hot path.')

  def fd_solver_forward(self, *args, **kwargs):
    raise NotImplementedError('PDE solvers are outside the B200 hot path.')
