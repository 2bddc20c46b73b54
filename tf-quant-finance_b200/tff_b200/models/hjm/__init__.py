"""HJM models and their Monte-Carlo pricers (`tf_quant_finance.models.hjm`)."""
from tff_b200.models.hjm.cap_floor import cap_floor_price
from tff_b200.models.hjm.gaussian_hjm import GaussianHJM
from tff_b200.models.hjm.quasi_gaussian_hjm import QuasiGaussianHJM
from tff_b200.models.hjm.swaption_pricing import price as swaption_price
from tff_b200.models.hjm.zero_coupon_bond_option import bond_option_price

__all__ = ['GaussianHJM', 'QuasiGaussianHJM', 'swaption_price', 'bond_option_price',
           'cap_floor_price']
