"""Mirror of `tf_quant_finance.math` restricted to the Monte-Carlo hot path."""
from tff_b200.math import piecewise
from tff_b200.math import qmc
from tff_b200.math import random
