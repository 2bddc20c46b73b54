"""Mirror of `tf_quant_finance.math.random` (`math/random_ops/__init__.py:17-32`)
for the generators on the Monte-Carlo hot path: Philox (PSEUDO / STATELESS and
their antithetic forms) and Sobol."""
from tff_b200.math.random import halton
from tff_b200.math.random import sobol
from tff_b200.math.random.multivariate_normal import multivariate_normal as mv_normal_sample
from tff_b200.math.random.multivariate_normal import RandomType
from tff_b200.math.random.philox import normal
from tff_b200.math.random.philox import stateless_normal
from tff_b200.math.random.philox import stateless_uniform
from tff_b200.math.random.stateless import stateless_random_shuffle
from tff_b200.math.random.uniform import uniform

__all__ = ['RandomType', 'mv_normal_sample', 'sobol', 'halton', 'stateless_normal',
           'normal', 'uniform', 'stateless_uniform', 'stateless_random_shuffle']
