"""Hull-White models and the Monte-Carlo swaption / bond option / cap pricers."""
from tff_b200.models.hull_white.cap_floor import cap_floor_price
from tff_b200.models.hull_white.one_factor import HullWhiteModel1F
from tff_b200.models.hull_white.swaption import bermudan_swaption_price
from tff_b200.models.hull_white.swaption import swaption_price
from tff_b200.models.hull_white.vector_hull_white import VectorHullWhiteModel
from tff_b200.models.hull_white.zero_coupon_bond_option import bond_option_price

__all__ = ['HullWhiteModel1F', 'VectorHullWhiteModel', 'swaption_price',
           'bermudan_swaption_price',
           'bond_option_price', 'cap_floor_price']
