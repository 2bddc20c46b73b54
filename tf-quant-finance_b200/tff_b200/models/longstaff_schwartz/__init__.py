"""Longstaff-Schwartz least-squares Monte Carlo."""
from tff_b200.models.longstaff_schwartz.lsm import least_square_mc
from tff_b200.models.longstaff_schwartz.lsm import make_polynomial_basis
from tff_b200.models.longstaff_schwartz.payoff_utils import make_basket_put_payoff
from tff_b200.models.longstaff_schwartz.payoff_utils import make_tabulated_payoff

__all__ = ['least_square_mc', 'make_polynomial_basis', 'make_basket_put_payoff',
           'make_tabulated_payoff']
