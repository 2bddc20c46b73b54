"""`tf_quant_finance.black_scholes` -- only the Brownian-bridge helpers that the
Monte-Carlo hot path uses for continuously monitored barriers; the
closed-form pricers of the package are outside the scope of the B200 engine."""
from tff_b200.black_scholes.brownian_bridge import brownian_bridge_double
from tff_b200.black_scholes.brownian_bridge import brownian_bridge_single

__all__ = ['brownian_bridge_double', 'brownian_bridge_single']
