"""Monte-Carlo interest-rate cap / floor pricing under one-factor Hull-White.

Drop-in for the simulation branch of
`tf_quant_finance.models.hull_white.cap_floor_price`
(`models/hull_white/cap_floor.py:30-235`): a caplet on the simple rate over
`[expiry, maturity]` struck at K is `(1 + tau K)` puts on the zero-coupon bond
P(expiry, maturity) struck at `1 / (1 + tau K)`; a floorlet the matching call.
All caplets of all caps share one fused simulation (one payoff slot each).
"""
import numpy as np

from tff_b200 import _tensor
from tff_b200.models.hull_white import zero_coupon_bond_option as zcb


def cap_floor_price(*,
                    strikes,
                    expiries,
                    maturities,
                    daycount_fractions,
                    reference_rate_fn,
                    mean_reversion,
                    volatility,
                    notional=1.0,
                    is_cap=True,
                    use_analytic_pricing=True,
                    num_samples=1,
                    random_type=None,
                    seed=None,
                    skip=0,
                    time_step=None,
                    dtype=None,
                    name=None):
  """Cap / floor prices of shape `strikes.shape[:-1]` (numpy array)."""
  del name
  dt_ = _tensor.infer_dtype(strikes, dtype, default=np.float32)
  strikes = _tensor.to_numpy(strikes, dt_)
  expiries = _tensor.to_numpy(expiries, dt_)
  maturities = _tensor.to_numpy(maturities, dt_)
  dcf = _tensor.to_numpy(daycount_fractions, dt_)
  notional = _tensor.to_numpy(notional, dt_)
  is_cap = np.asarray(_tensor.to_numpy(is_cap), dtype=bool)
  bond_option_strikes = (1.0 / (1.0 + dcf * strikes)).astype(dt_)
  caplet_prices = zcb.bond_option_price(
      strikes=bond_option_strikes,
      expiries=expiries,
      maturities=maturities,
      discount_rate_fn=reference_rate_fn,
      mean_reversion=mean_reversion,
      volatility=volatility,
      is_call_options=~is_cap,
      use_analytic_pricing=use_analytic_pricing,
      num_samples=num_samples,
      random_type=random_type,
      seed=seed,
      skip=skip,
      time_step=time_step,
      dtype=dt_)
  caplet_prices = np.where(np.broadcast_to(expiries, caplet_prices.shape) < 0.0,
                           np.zeros_like(caplet_prices), caplet_prices)
  return np.sum(notional * (1.0 + dcf * strikes) * caplet_prices, axis=-1).astype(dt_)
