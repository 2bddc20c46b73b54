"""Monte-Carlo interest-rate cap / floor price under the HJM model
(`models/hjm/cap_floor.py:27-229`): a caplet on the simple rate over
`[expiry, maturity]` struck at K is `(1 + tau K)` puts on the zero-coupon bond
P(expiry, maturity) struck at `1 / (1 + tau K)`, a floorlet the matching call; all
caplets of all caps share one fused simulation (one payoff slot each)."""
import numpy as np

from tff_b200 import _tensor
from tff_b200.models.hjm import zero_coupon_bond_option as zcb


def cap_floor_price(*, strikes, expiries, maturities, daycount_fractions, reference_rate_fn, dim,
                    mean_reversion, volatility, corr_matrix=None, notional=1.0, is_cap=True,
                    num_samples=1, random_type=None, seed=None, skip=0, time_step=None, dtype=None,
                    name=None):
  """`tff.models.hjm.cap_floor_price`: prices of shape `strikes.shape[:-1]` (numpy)."""
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
      strikes=bond_option_strikes, expiries=expiries, maturities=maturities,
      discount_rate_fn=reference_rate_fn, dim=dim, mean_reversion=mean_reversion,
      volatility=volatility, corr_matrix=corr_matrix, is_call_options=~is_cap,
      num_samples=num_samples, random_type=random_type, seed=seed, skip=skip, time_step=time_step,
      dtype=dt_)
  caplet_prices = np.where(np.broadcast_to(expiries, caplet_prices.shape) < 0.0,
                           np.zeros_like(caplet_prices), caplet_prices)
  return np.sum(notional * (1.0 + dcf * strikes) * caplet_prices, axis=-1).astype(dt_)
