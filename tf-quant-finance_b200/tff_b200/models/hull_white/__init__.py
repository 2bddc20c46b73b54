"""Hull-White model and the Monte-Carlo swaption pricer."""
from tff_b200.models.hull_white.one_factor import HullWhiteModel1F
from tff_b200.models.hull_white.swaption import swaption_price

__all__ = ['HullWhiteModel1F', 'swaption_price']
