"""tff_b200: B200-native Monte-Carlo path engine behind tf-quant-finance's
Euler sampling API.

    import tff_b200 as tff
    process = tff.models.HestonModel(...)
    paths = process.sample_paths(times, num_samples=..., random_type=tff.math.random.RandomType.SOBOL, ...)

Only the hot path of the reference is mirrored:
`tff.math.random`, `tff.models.euler_sampling`, `GenericItoProcess`, the GBM /
Heston / Hull-White model classes, `swaption_price` and Longstaff-Schwartz.
All device work runs in libtqf.so (hand-written CUDA for sm_100a); there is no
CPU fallback.
"""
from tff_b200 import black_scholes
from tff_b200 import math
from tff_b200 import models

__all__ = ['black_scholes', 'math', 'models']
