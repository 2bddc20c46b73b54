"""Geometric Brownian motion models."""
from tff_b200.models.geometric_brownian_motion.univariate_geometric_brownian_motion import GeometricBrownianMotion

__all__ = ['GeometricBrownianMotion']
