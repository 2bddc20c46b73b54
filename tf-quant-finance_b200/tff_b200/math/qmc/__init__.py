"""Mirror of `tf_quant_finance.math.qmc` (`math/qmc/__init__.py:17-38`):
digital nets, Sobol generating matrices and lattice rules, sampled on the
device by libtqf (`csrc/tqf_qmc.cu`)."""
from tff_b200.math.qmc import utils
from tff_b200.math.qmc.digital_net import digital_net_sample
from tff_b200.math.qmc.digital_net import random_digital_shift
from tff_b200.math.qmc.digital_net import random_scrambling_matrices
from tff_b200.math.qmc.digital_net import scramble_generating_matrices
from tff_b200.math.qmc.lattice_rule import lattice_rule_sample
from tff_b200.math.qmc.lattice_rule import random_scrambling_vectors
from tff_b200.math.qmc.sobol import sobol_generating_matrices
from tff_b200.math.qmc.sobol import sobol_sample

__all__ = [
    'digital_net_sample',
    'lattice_rule_sample',
    'random_digital_shift',
    'random_scrambling_matrices',
    'random_scrambling_vectors',
    'scramble_generating_matrices',
    'sobol_generating_matrices',
    'sobol_sample',
    'utils',
]
