"""Mirror of `tf_quant_finance.models` for the Monte-Carlo hot path."""
from tff_b200.models import closures
from tff_b200.models import euler_sampling
from tff_b200.models import hjm
from tff_b200.models import hull_white
from tff_b200.models import longstaff_schwartz
from tff_b200.models import milstein_sampling
from tff_b200.models import utils
from tff_b200.models.generic_ito_process import GenericItoProcess
from tff_b200.models.geometric_brownian_motion.multivariate_geometric_brownian_motion import MultivariateGeometricBrownianMotion
from tff_b200.models.geometric_brownian_motion.univariate_geometric_brownian_motion import GeometricBrownianMotion
from tff_b200.models.heston.heston_model import HestonModel
from tff_b200.models.hull_white.one_factor import HullWhiteModel1F
from tff_b200.models.hull_white.vector_hull_white import VectorHullWhiteModel
from tff_b200.models.ito_process import ItoProcess

__all__ = ['closures', 'euler_sampling', 'milstein_sampling', 'utils', 'GenericItoProcess',
           'GeometricBrownianMotion', 'MultivariateGeometricBrownianMotion', 'HestonModel', 'HullWhiteModel1F', 'VectorHullWhiteModel', 'ItoProcess',
           'hull_white', 'hjm', 'longstaff_schwartz']
