"""Heston stochastic volatility model."""
from tff_b200.models.heston.heston_model import HestonModel

__all__ = ['HestonModel']
