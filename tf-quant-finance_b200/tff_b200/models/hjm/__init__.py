"""HJM models and the Monte-Carlo swaption pricer (`tf_quant_finance.models.hjm`)."""
from tff_b200.models.hjm.gaussian_hjm import GaussianHJM
from tff_b200.models.hjm.quasi_gaussian_hjm import QuasiGaussianHJM
from tff_b200.models.hjm.swaption_pricing import price as swaption_price

__all__ = ['GaussianHJM', 'QuasiGaussianHJM', 'swaption_price']
